"""SURVEY 8(f) items 3 and 4 without SimpleITK / nibabel:

* the reference's own end-to-end fixture test (tests/test_convex_adam_mind.py:45-85 there): T2w and ADC prostate MR volumes
  read from MetaImage files, resampled to 1 mm on the fixed image's grid (convexadam_b200/lite_io.py restates
  `resample_img` / `resample_moving_to_fixed`), registered with the package defaults, warped with apply_convex and
  compared with the reference's golden warped image by 3-D SSIM > 0.95.  The image fixtures are the reference's (staged
  byte for byte under the git-ignored oracle/_ref/tests by oracle/stage_reference.py); skipped where they are absent.
  CPU: through the oracle; GPU: through convex_adam_pt / apply_convex.
* `disp.nii.gz`: the NIfTI-1 writer / reader behind `convex_adam(...)` when nibabel is not installed -- header fields
  against the NIfTI-1 specification, data and affine round trip, Fortran data order.
"""
import gzip
import os
import struct

import numpy as np
import pytest
import torch

from convexadam_b200 import lite_io as L
from tests.ssim import ssim3d

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def fixture_dir():
    for base in (os.environ.get("CONVEXADAM_REFERENCE"), "/root/reference", os.path.join(ROOT, "oracle", "_ref")):
        if base and os.path.exists(os.path.join(base, "tests", "input", "10000", "10000_1000000_t2w.mha")):
            return os.path.join(base, "tests")
    return None


@pytest.fixture(scope="module")
def prostate_case():
    d = fixture_dir()
    if d is None:
        pytest.skip("reference image fixtures not staged (oracle/stage_reference.py)")
    t2w = L.read_mha(os.path.join(d, "input", "10000", "10000_1000000_t2w.mha"))
    adc = L.read_mha(os.path.join(d, "input", "10000", "10000_1000000_adc.mha"))
    golden = L.read_mha(os.path.join(d, "output-expected", "10000", "10000_1000000_adc_warped.mha"))
    fixed = L.resample_img(t2w, spacing=(1.0, 1.0, 1.0))
    moving = L.resample_moving_to_fixed(fixed, adc)
    assert fixed.array.shape == golden.array.shape == moving.array.shape == (102, 180, 180)
    assert np.allclose(fixed.direction, golden.direction) and np.allclose(fixed.origin, golden.origin)
    return fixed, moving, golden


def test_fixture_registration_oracle(prostate_case):
    from oracle import ca_oracle as orc
    fixed, moving, golden = prostate_case
    disp = orc.convex_adam_pt(fixed.array.astype(np.float32), moving.array.astype(np.float32))        # package defaults
    warped = orc.apply_convex(disp, moving.array.astype(np.float64)).astype(np.float32)
    s, s0 = ssim3d(warped, golden.array.copy()), ssim3d(moving.array.astype(np.float32), golden.array.copy())
    print(f"[oracle] fixture test: SSIM(warped, reference's golden) = {s:.4f} (unregistered moving image: {s0:.4f})")
    assert s > 0.95 and s0 < 0.7


@pytest.mark.gpu
def test_fixture_registration_cuda(prostate_case, tmp_path):
    from convexadam_b200.apply_convex import apply_convex
    from convexadam_b200.convex_adam_MIND import convex_adam_pt
    fixed, moving, golden = prostate_case
    disp = convex_adam_pt(img_fixed=fixed.array, img_moving=moving.array)      # as the reference's test: all defaults
    warped = apply_convex(disp=disp, moving=moving.array)
    assert warped.dtype == moving.array.dtype
    out = L.Image(warped.astype(np.float32), moving.spacing, moving.origin, moving.direction)
    L.write_mha(tmp_path / "adc_warped.mha", out)
    back = L.read_mha(tmp_path / "adc_warped.mha")
    s = ssim3d(back.array.copy(), golden.array.copy())
    print(f"[cuda] fixture test: SSIM(warped, reference's golden) = {s:.4f}")
    assert s > 0.95
    # identity case of the same test file (tests/test_convex_adam_mind.py:22-42): field within 0.1 voxel of zero
    ident = convex_adam_pt(img_fixed=fixed.array, img_moving=fixed.array)
    assert np.allclose(ident, 0.0, atol=0.1)


def test_mha_round_trip_and_geometry(tmp_path):
    rs = np.random.RandomState(0)
    ang = 0.4
    direction = np.array([[1, 0, 0], [0, np.cos(ang), np.sin(ang)], [0, -np.sin(ang), np.cos(ang)]])
    for dt, comp in ((np.uint16, True), (np.float32, False), (np.int16, True)):
        img = L.Image((rs.rand(5, 7, 9) * 1000).astype(dt), (0.5, 0.75, 3.3), (-10.0, 4.0, 7.5), direction)
        L.write_mha(tmp_path / "a.mha", img, compressed=comp)
        back = L.read_mha(tmp_path / "a.mha")
        assert back.array.dtype == dt and (back.array == img.array).all()
        assert np.allclose(back.spacing, img.spacing) and np.allclose(back.origin, img.origin) and np.allclose(back.direction, direction)
    # resampling onto the image's own grid is the identity; a shifted grid samples between the voxels
    img = L.Image(rs.rand(6, 7, 8).astype(np.float32), (1.0, 2.0, 3.0), (1.0, -2.0, 0.5), direction)
    same = L.resample_to_grid(img, img.size, img.spacing, img.origin, img.direction)
    assert np.allclose(same.array, img.array, atol=1e-6)
    half = L.resample_to_grid(img, img.size, img.spacing, tuple(np.asarray(img.origin) + direction @ np.array([0.5, 0.0, 0.0])), img.direction)
    assert np.allclose(half.array[:, :, :-1], 0.5 * (img.array[:, :, :-1] + img.array[:, :, 1:]), atol=1e-6)
    assert np.allclose(half.array[:, :, -1], 0.0)            # ITK: at the last sample + 0.5 a point is outside the buffer
    # the reference's size rule: int(size * old / new + 0.5)
    assert L.resample_img(img, (2.0, 2.0, 2.0)).size == (4, 7, 9)      # (8*1, 7*2, 6*3) mm / 2 mm


def test_nifti_writer_header_and_round_trip(tmp_path):
    rs = np.random.RandomState(1)
    disp = rs.randn(6, 5, 4, 3)                               # (H,W,D,3) float64, the layout convex_adam writes
    affine = np.array([[0.0, -2.0, 0.0, 10.0], [1.5, 0.0, 0.0, -3.0], [0.0, 0.0, 3.0, 7.0], [0.0, 0.0, 0.0, 1.0]])
    path = tmp_path / "disp.nii.gz"
    L.write_nifti(path, disp, affine)
    with gzip.open(path, "rb") as f:
        blob = f.read()
    # NIfTI-1 specification (nifti1.h): header size, dim, datatype / bitpix, vox_offset, magic, sform
    assert struct.unpack_from("<i", blob, 0)[0] == 348
    assert struct.unpack_from("<8h", blob, 40) == (4, 6, 5, 4, 3, 1, 1, 1)
    assert struct.unpack_from("<2h", blob, 70) == (64, 64)                     # NIFTI_TYPE_FLOAT64
    assert struct.unpack_from("<f", blob, 108)[0] == 352.0 and blob[344:348] == b"n+1\x00"
    assert struct.unpack_from("<2h", blob, 252) == (0, 2)                      # qform unknown, sform aligned (nibabel's choice)
    assert np.allclose(struct.unpack_from("<4f", blob, 76)[1:], [1.5, 2.0, 3.0])   # pixdim = column norms of the affine
    assert np.allclose(np.array(struct.unpack_from("<12f", blob, 280)).reshape(3, 4), affine[:3])
    assert len(blob) == 352 + disp.size * 8
    first = np.frombuffer(blob, np.float64, count=6, offset=352)                # Fortran order: the first axis runs fastest
    assert (first == disp[:, 0, 0, 0]).all()
    data, aff = L.read_nifti(path)
    assert data.dtype == np.float64 and (data == disp).all() and np.allclose(aff, affine)
    # the quaternion stored next to the sform reproduces the affine's rotation (qform readers)
    b, c, d = struct.unpack_from("<3f", blob, 256)
    a = np.sqrt(1.0 - b * b - c * c - d * d)
    r = np.array([[a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c)],
                  [2 * (b * c + a * d), a * a + c * c - b * b - d * d, 2 * (c * d - a * b)],
                  [2 * (b * d - a * c), 2 * (c * d + a * b), a * a + d * d - b * b - c * c]])
    qfac = struct.unpack_from("<f", blob, 76)[0]
    assert np.allclose(r @ np.diag([1.5, 2.0, 3.0 * qfac]), affine[:3, :3], atol=1e-5)


@pytest.mark.gpu
def test_convex_adam_file_wrapper_without_nibabel(tmp_path):
    """convex_adam(path_fixed, path_moving, ..., result_path): NIfTI in, disp.nii.gz out with the fixed image's affine and
    convex_adam_pt's field (reference MIND.py:205-248)."""
    from convexadam_b200.convex_adam_MIND import convex_adam, convex_adam_pt
    from convexadam_b200.synth import make_pair
    fixed, moving = make_pair((40, 36, 44), seed=9, disp_scale=0.1)
    affine = np.diag([1.5, 1.5, 2.0, 1.0])
    affine[:3, 3] = (-20.0, 11.0, 3.0)
    L.write_nifti(tmp_path / "fixed.nii.gz", fixed.numpy(), affine)
    L.write_nifti(tmp_path / "moving.nii.gz", moving.numpy(), affine)
    kw = dict(grid_sp=3, disp_hw=3, selected_niter=6)
    convex_adam(str(tmp_path / "fixed.nii.gz"), str(tmp_path / "moving.nii.gz"), result_path=str(tmp_path), **kw)
    data, aff = L.read_nifti(tmp_path / "disp.nii.gz")
    want = convex_adam_pt(fixed, moving, **kw)
    assert data.shape == (40, 36, 44, 3) and data.dtype == np.float64 and (data == want).all() and np.allclose(aff, affine)
