"""Drop-in for the reference's src/convexAdam/convex_adam_MIND.py: same entry points (extract_features :22,
convex_adam_pt :64, convex_adam :205, CLI :251), same defaults, same returned layout -- (H,W,D,3) float64,
[..., c] = displacement in voxels along array axis c -- computed by the B200 kernels of libconvexadam_b200.
"""
import argparse
import os
import time
import warnings
from pathlib import Path
from typing import Optional, Union

import numpy as np
import torch

from . import pipeline
from ._native import get_lib
from .convex_adam_utils import MINDSSC, validate_image

warnings.filterwarnings("ignore")


def _default_device():
    return torch.device("cuda" if torch.cuda.is_available() else "cpu")


def _require_cuda(device):
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("convexadam_b200 runs on CUDA devices only (B200, sm_100a); there is no CPU path. "
                           "Use the reference package for CPU execution.")
    return device


def _fill_outside_mask(img, mask, device):
    """use_mask branch of the reference (MIND.py:36-54): erode the mask (`avg3(mask) > 0.9`), give every voxel of the 2x
    subsampled grid the intensity of the nearest voxel inside the eroded mask (the reference: scipy's distance transform
    on the host; here ca_nearest_inside, the same feature transform on the device, equidistant choices included),
    upsample trilinearly and keep the original intensities inside the mask.  Like the reference this needs even H, W, D."""
    lib = get_lib()
    H, W, D = img.shape[-3:]
    img = img.reshape(H, W, D).to(device).float().contiguous()
    m = lib.mask_erode(mask.reshape(H, W, D).to(device).float().contiguous())
    m_sub = m[::2, ::2, ::2].contiguous()
    idx = lib.nearest_inside(m_sub).long()
    h2, w2, d2 = m_sub.shape
    # the reference's flat index into the subsampled image: idx[0]*D//2*W//2 + idx[1]*D//2 + idx[2]  (MIND.py:44)
    flat = idx[0] * D // 2 * W // 2 + idx[1] * D // 2 + idx[2]
    sub = img[::2, ::2, ::2].reshape(-1)[flat].reshape(1, h2, w2, d2).contiguous()
    filled = lib.trilinear(sub, (2 * h2, 2 * w2, 2 * d2))[0]                 # F.interpolate(scale_factor=2, mode='trilinear')
    if filled.shape != img.shape:
        raise RuntimeError(f"use_mask needs even image dimensions (got {tuple(img.shape)}): the reference fails on the same line")
    return torch.where(m != 0, img, filled).contiguous()


def extract_features(img_fixed, img_moving, mind_r, mind_d, use_mask, mask_fixed, mask_moving,
                     device=torch.device("cuda"), dtype=torch.float16):
    """MIND-SSC features of both images, (1,12,H,W,D) each, cast to `dtype` (reference MIND.py:22-61)."""
    device = _require_cuda(device)
    if use_mask:
        img_fixed = _fill_outside_mask(img_fixed, mask_fixed, device)
        img_moving = _fill_outside_mask(img_moving, mask_moving, device)
    ff = MINDSSC(img_fixed.reshape((1, 1) + tuple(img_fixed.shape[-3:])).to(device), mind_r, mind_d, device=device).to(dtype)
    fm = MINDSSC(img_moving.reshape((1, 1) + tuple(img_moving.shape[-3:])).to(device), mind_r, mind_d, device=device).to(dtype)
    return ff, fm


_COPY_STREAMS = {}
_FP16_WARNED = False


def _warn_fp16_once():
    """The reference silently switches arithmetic with `dtype`; this package computes in fp32 either way and says so once.
    (Issued through `warnings` with an explicit filter entry: the reference's blanket `filterwarnings("ignore")`, which
    this module repeats at import, would otherwise swallow it.)"""
    global _FP16_WARNED
    if _FP16_WARNED:
        return
    _FP16_WARNED = True
    with warnings.catch_warnings():
        warnings.simplefilter("always")
        warnings.warn("convexadam_b200: dtype=torch.float16 computes features and cost volume in float32 and rounds only the returned "
                      "field through half (MIND.py:198-200); the reference's half-precision cost volume is not reproduced. "
                      "Pass dtype=torch.float32 for the parity-pinned path.", RuntimeWarning, stacklevel=3)


def _copy_stream(device):
    key = str(device)
    if key not in _COPY_STREAMS:
        _COPY_STREAMS[key] = torch.cuda.Stream(device)
    return _COPY_STREAMS[key]


def convex_adam_pt(
    img_fixed,
    img_moving,
    mind_r: int = 1,
    mind_d: int = 2,
    lambda_weight: float = 1.25,
    grid_sp: int = 6,
    disp_hw: int = 4,
    selected_niter: int = 80,
    selected_smooth: int = 0,
    grid_sp_adam: int = 2,
    ic: bool = True,
    use_mask: bool = False,
    path_fixed_mask: Optional[Union[Path, str]] = None,
    path_moving_mask: Optional[Union[Path, str]] = None,
    dtype: torch.dtype = torch.float16,
    verbose: bool = False,
    device: torch.device = None,
    split_over=None,
) -> np.ndarray:
    """Coupled convex optimisation with adam instance optimisation (reference MIND.py:64-202).

    split_over (extension, not in the reference): a torch.distributed process group (or True for the default
    group) over which this ONE pair is split into H slabs -- every rank calls with the same images and gets the
    same field (convexadam_b200/slab.py; for pairs whose cost volume does not fit one GPU).

    dtype: float32 computes and returns fp32 values (the reference's CPU policy, parity-pinned against it);
    float16 (the reference's GPU default) keeps features and cost volume in fp32 and rounds the returned field
    through half exactly where the reference does (MIND.py:198-200): it does NOT reproduce the reference's half
    arithmetic (MIND.py:58-59,118-119; utils.py:84-85,98-104), whose selections differ from the fp32 ones.  A
    one-time warning says so; tests/test_reference_cuda_gpu.py measures the distance to the reference's own CUDA fp16 run."""
    device = _require_cuda(device if device is not None else _default_device())
    lib = get_lib()
    if dtype == torch.float16:
        _warn_fp16_once()
    img_fixed = validate_image(img_fixed).float()
    img_moving = validate_image(img_moving).float()
    if dtype not in (torch.float16, torch.float32):
        raise ValueError("dtype must be torch.float16 or torch.float32")
    if img_fixed.dim() != 3 or img_fixed.shape != img_moving.shape:
        raise ValueError("img_fixed and img_moving must be 3-D volumes of the same shape")
    prm = pipeline.Params(mind_r, mind_d, lambda_weight, grid_sp, disp_hw, selected_niter, selected_smooth, grid_sp_adam, ic)

    t0 = time.time()
    with torch.cuda.device(device):
        fixed_dev = img_fixed.to(device, non_blocking=True).contiguous()
        moving_ready = None
        if img_moving.is_cuda or use_mask or (split_over is not None and split_over is not False):
            moving_dev = img_moving.to(device, non_blocking=True).contiguous()
        else:
            # the moving image travels on a side stream while the fixed image's descriptors are computed
            main = torch.cuda.current_stream(device)
            side = _copy_stream(device)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                moving_dev = img_moving.to(device, non_blocking=True).contiguous()
            moving_ready = side.record_event()
            moving_dev.record_stream(main)
        if use_mask:
            load, _ = _nifti_io()
            mask_fixed = torch.from_numpy(np.asarray(load(path_fixed_mask)[0])).float()
            mask_moving = torch.from_numpy(np.asarray(load(path_moving_mask)[0])).float()
            fixed_dev = _fill_outside_mask(fixed_dev, mask_fixed, device)
            moving_dev = _fill_outside_mask(moving_dev, mask_moving, device)
        if split_over is not None and split_over is not False:
            from . import slab
            disp = slab.register_mind_slab(fixed_dev, moving_dev, prm, lib, None if split_over is True else split_over)
        else:
            disp = pipeline.register_mind(fixed_dev, moving_dev, prm, lib, moving_ready=moving_ready)
        packed = lib.pack_field(disp, round_half=(dtype == torch.float16))
        host = torch.empty(packed.shape, dtype=torch.float64, pin_memory=True)
        host.copy_(packed, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    if verbose:
        print(f"case time: {time.time() - t0}")
    return host.numpy()


def convex_adam(
    path_img_fixed,
    path_img_moving,
    mind_r: int = 1,
    mind_d: int = 2,
    lambda_weight: float = 1.25,
    grid_sp: int = 6,
    disp_hw: int = 4,
    selected_niter: int = 80,
    selected_smooth: int = 0,
    grid_sp_adam: int = 2,
    ic: bool = True,
    use_mask: bool = False,
    path_fixed_mask=None,
    path_moving_mask=None,
    result_path="./",
    verbose: bool = False,
) -> None:
    """File-in / disp.nii.gz-out wrapper (reference MIND.py:205-248): NIfTI images in, `result_path/disp.nii.gz` out with
    the fixed image's affine.  Uses nibabel when it is installed, else this package's own NIfTI-1 reader / writer."""
    load, save = _nifti_io()
    fixed_data, affine = load(path_img_fixed)
    img_fixed = torch.from_numpy(np.asarray(fixed_data)).float()
    img_moving = torch.from_numpy(np.asarray(load(path_img_moving)[0])).float()
    displacements = convex_adam_pt(img_fixed=img_fixed, img_moving=img_moving, mind_r=mind_r, mind_d=mind_d,
                                   lambda_weight=lambda_weight, grid_sp=grid_sp, disp_hw=disp_hw,
                                   selected_niter=selected_niter, selected_smooth=selected_smooth,
                                   grid_sp_adam=grid_sp_adam, ic=ic, use_mask=use_mask, path_fixed_mask=path_fixed_mask,
                                   path_moving_mask=path_moving_mask, verbose=verbose)
    save(os.path.join(result_path, "disp.nii.gz"), displacements, affine)


def _nifti_io():
    """(load(path) -> (float64 data, affine), save(path, data, affine)): nibabel if present, else convexadam_b200.lite_io."""
    try:
        import nibabel as nib

        def load(path):
            img = nib.load(path)
            return img.get_fdata(), img.affine

        def save(path, data, affine):
            nib.save(nib.Nifti1Image(data, affine), path)
    except ImportError:
        from . import lite_io

        def load(path):
            data, affine = lite_io.read_nifti(path)
            return np.asarray(data, np.float64), affine

        save = lite_io.write_nifti
    return load, save


def _main():
    parser = argparse.ArgumentParser()
    parser.add_argument("-f", "--path_img_fixed", type=str, required=True)
    parser.add_argument("-m", "--path_img_moving", type=str, required=True)
    parser.add_argument("--mind_r", type=int, default=1)
    parser.add_argument("--mind_d", type=int, default=2)
    parser.add_argument("--lambda_weight", type=float, default=1.25)
    parser.add_argument("--grid_sp", type=int, default=6)
    parser.add_argument("--disp_hw", type=int, default=4)
    parser.add_argument("--selected_niter", type=int, default=80)
    parser.add_argument("--selected_smooth", type=int, default=0)
    parser.add_argument("--grid_sp_adam", type=int, default=2)
    parser.add_argument("--ic", choices=("True", "False"), default="True")
    parser.add_argument("--use_mask", choices=("True", "False"), default="False")
    parser.add_argument("--path_mask_fixed", type=str, default=None)
    parser.add_argument("--path_mask_moving", type=str, default=None)
    parser.add_argument("--result_path", type=str, default="./")
    a = parser.parse_args()
    convex_adam(a.path_img_fixed, a.path_img_moving, a.mind_r, a.mind_d, a.lambda_weight, a.grid_sp, a.disp_hw,
                a.selected_niter, a.selected_smooth, a.grid_sp_adam, a.ic == "True", a.use_mask == "True",
                a.path_mask_fixed, a.path_mask_moving, a.result_path)


if __name__ == "__main__":
    _main()
