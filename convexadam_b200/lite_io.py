"""Dependency-free medical-image I/O next to the hot path: MetaImage (.mha) and NIfTI-1 (.nii / .nii.gz) read / write and
linear resampling onto a grid, in numpy.

The reference reads and writes through SimpleITK and nibabel (tests/test_convex_adam_mind.py:45-85 resamples the inputs
with `resample_img` / `resample_moving_to_fixed`, convex_adam_utils.py:282-306; `convex_adam` writes `disp.nii.gz` with
nibabel, convex_adam_MIND.py:246-248).  Neither library exists in this image, so the file-facing wrappers of this package
fall back to this module.  Geometry follows ITK: `index -> physical = origin +
direction @ (spacing * index)`, arrays are indexed (z, y, x) like `sitk.GetArrayFromImage`.
"""
import gzip
import os
import struct
import zlib
from dataclasses import dataclass, field

import numpy as np

_MET = {"MET_UCHAR": np.uint8, "MET_CHAR": np.int8, "MET_USHORT": np.uint16, "MET_SHORT": np.int16, "MET_UINT": np.uint32,
        "MET_INT": np.int32, "MET_ULONG": np.uint64, "MET_LONG": np.int64, "MET_FLOAT": np.float32, "MET_DOUBLE": np.float64}
_MET_INV = {np.dtype(v): k for k, v in _MET.items()}


@dataclass
class Image:
    """A 3-D image with ITK geometry.  `array` is (z, y, x); `spacing`, `origin` are (x, y, z); `direction` is the 3x3
    matrix whose COLUMNS are the physical directions of the index axes x, y, z (ITK's convention)."""
    array: np.ndarray
    spacing: tuple = (1.0, 1.0, 1.0)
    origin: tuple = (0.0, 0.0, 0.0)
    direction: np.ndarray = field(default_factory=lambda: np.eye(3))

    @property
    def size(self):                      # (x, y, z), as sitk.Image.GetSize()
        return tuple(int(s) for s in self.array.shape[::-1])

    def index_to_physical_matrix(self):
        """4x4 affine: homogeneous (x, y, z) index -> physical point."""
        a = np.eye(4)
        a[:3, :3] = np.asarray(self.direction, float) @ np.diag(self.spacing)
        a[:3, 3] = self.origin
        return a


# ---- MetaImage ------------------------------------------------------------------------------------------------------------
def read_mha(path) -> Image:
    """Single-file MetaImage (`ElementDataFile = LOCAL`), raw or zlib-compressed, little endian, scalar pixels."""
    with open(path, "rb") as f:
        blob = f.read()
    hdr, pos = {}, 0
    while True:
        end = blob.index(b"\n", pos)
        line = blob[pos:end].decode("latin-1").strip()
        pos = end + 1
        if "=" in line:
            k, v = (s.strip() for s in line.split("=", 1))
            hdr[k] = v
            if k == "ElementDataFile":
                break
    if hdr.get("ElementDataFile") != "LOCAL" or int(hdr.get("NDims", 3)) != 3:
        raise ValueError(f"{path}: only single-file 3-D MetaImages are supported")
    if hdr.get("BinaryDataByteOrderMSB", hdr.get("ElementByteOrderMSB", "False")).lower() == "true":
        raise ValueError(f"{path}: big-endian data not supported")
    if int(hdr.get("ElementNumberOfChannels", 1)) != 1:
        raise ValueError(f"{path}: vector pixels not supported")
    dtype = np.dtype(_MET[hdr["ElementType"]])
    nx, ny, nz = (int(v) for v in hdr["DimSize"].split())
    data = blob[pos:]
    if hdr.get("CompressedData", "False").lower() == "true":
        data = zlib.decompress(data)
    arr = np.frombuffer(data, dtype=dtype, count=nx * ny * nz).reshape(nz, ny, nx)
    spacing = tuple(float(v) for v in hdr.get("ElementSpacing", "1 1 1").split())
    origin = tuple(float(v) for v in hdr.get("Offset", hdr.get("Position", "0 0 0")).split())
    tm = np.array([float(v) for v in hdr.get("TransformMatrix", "1 0 0 0 1 0 0 0 1").split()]).reshape(3, 3)
    # MetaIO stores the direction cosines of index axis i as the i-th ROW of TransformMatrix; ITK's direction matrix holds
    # them as columns
    return Image(arr, spacing, origin, tm.T.copy())


def write_mha(path, img: Image, compressed=False):
    arr = np.ascontiguousarray(img.array)
    nz, ny, nx = arr.shape
    data = arr.tobytes()
    lines = ["ObjectType = Image", "NDims = 3", "BinaryData = True", "BinaryDataByteOrderMSB = False",
             f"CompressedData = {'True' if compressed else 'False'}"]
    if compressed:
        data = zlib.compress(data)
        lines.append(f"CompressedDataSize = {len(data)}")
    tm = np.asarray(img.direction, float).T.reshape(-1)
    lines += ["TransformMatrix = " + " ".join(repr(float(v)) for v in tm), "Offset = " + " ".join(repr(float(v)) for v in img.origin),
              "CenterOfRotation = 0 0 0", "AnatomicalOrientation = RAI", "ElementSpacing = " + " ".join(repr(float(v)) for v in img.spacing),
              f"DimSize = {nx} {ny} {nz}", f"ElementType = {_MET_INV[arr.dtype]}", "ElementDataFile = LOCAL"]
    with open(path, "wb") as f:
        f.write(("\n".join(lines) + "\n").encode("latin-1"))
        f.write(data)


# ---- resampling (ITK ResampleImageFilter with an identity transform and linear interpolation) ------------------------------
def resample_to_grid(img: Image, size, spacing, origin, direction, default=0.0) -> Image:
    """`img` sampled on the grid (size (x,y,z), spacing, origin, direction): every output voxel's physical point is mapped
    to a continuous index of `img` and interpolated linearly; ITK's rules at the border (a point within half a voxel
    outside the first / last sample takes that sample, anything further out takes `default`)."""
    nx, ny, nz = (int(s) for s in size)
    out_aff = np.eye(4)
    out_aff[:3, :3] = np.asarray(direction, float) @ np.diag(spacing)
    out_aff[:3, 3] = origin
    m = np.linalg.inv(img.index_to_physical_matrix()) @ out_aff            # output index -> input continuous index
    src = img.array.astype(np.float64)
    sz, sy, sx = src.shape
    out = np.empty((nz, ny, nx), np.float64)
    xs = np.arange(nx, dtype=np.float64)
    ys = np.arange(ny, dtype=np.float64)
    for k in range(nz):                                                  # plane by plane: bounded memory
        cx = m[0, 0] * xs[None, :] + m[0, 1] * ys[:, None] + m[0, 2] * k + m[0, 3]
        cy = m[1, 0] * xs[None, :] + m[1, 1] * ys[:, None] + m[1, 2] * k + m[1, 3]
        cz = m[2, 0] * xs[None, :] + m[2, 1] * ys[:, None] + m[2, 2] * k + m[2, 3]
        inside = (cx >= -0.5) & (cx < sx - 0.5) & (cy >= -0.5) & (cy < sy - 0.5) & (cz >= -0.5) & (cz < sz - 0.5)
        val = np.zeros((ny, nx), np.float64)
        c = [np.clip(cx, 0.0, sx - 1.0), np.clip(cy, 0.0, sy - 1.0), np.clip(cz, 0.0, sz - 1.0)]
        f0 = [np.floor(v) for v in c]
        t = [c[i] - f0[i] for i in range(3)]
        i0 = [f0[i].astype(np.int64) for i in range(3)]
        i1 = [np.minimum(i0[0] + 1, sx - 1), np.minimum(i0[1] + 1, sy - 1), np.minimum(i0[2] + 1, sz - 1)]
        for dz in (0, 1):
            wz = t[2] if dz else 1.0 - t[2]
            zz = i1[2] if dz else i0[2]
            for dy in (0, 1):
                wy = t[1] if dy else 1.0 - t[1]
                yy = i1[1] if dy else i0[1]
                for dx in (0, 1):
                    wx = t[0] if dx else 1.0 - t[0]
                    xx = i1[0] if dx else i0[0]
                    val += src[zz, yy, xx] * (wz * wy * wx)
        out[k] = np.where(inside, val, default)
    # like ITK's ResampleImageFilter the output keeps the input's pixel type (integers: rounded, saturated)
    if img.array.dtype.kind in "iu":
        info = np.iinfo(img.array.dtype)
        out = np.clip(np.rint(out), info.min, info.max)
    return Image(out.astype(img.array.dtype), tuple(spacing), tuple(origin), np.asarray(direction, float).copy())


def resample_img(img: Image, spacing=(1.0, 1.0, 1.0)) -> Image:
    """The reference's `resample_img` (convex_adam_utils.py:282-292): same origin and direction, new spacing, size
    int(size * old / new + 0.5)."""
    size = [int(sz * spc / new + 0.5) for sz, spc, new in zip(img.size, img.spacing, spacing)]
    return resample_to_grid(img, size, spacing, img.origin, img.direction)


def resample_moving_to_fixed(fixed: Image, moving: Image) -> Image:
    """The reference's `resample_moving_to_fixed` (convex_adam_utils.py:295-306)."""
    return resample_to_grid(moving, fixed.size, fixed.spacing, fixed.origin, fixed.direction)


# ---- NIfTI-1 --------------------------------------------------------------------------------------------------------------
_NIFTI_DT = {np.dtype(np.uint8): 2, np.dtype(np.int16): 4, np.dtype(np.int32): 8, np.dtype(np.float32): 16, np.dtype(np.float64): 64,
             np.dtype(np.int8): 256, np.dtype(np.uint16): 512, np.dtype(np.uint32): 768, np.dtype(np.int64): 1024, np.dtype(np.uint64): 1280}
_NIFTI_DT_INV = {v: k for k, v in _NIFTI_DT.items()}


def _quaternion_from_affine(affine):
    """NIfTI qform parameters (b, c, d, qfac, pixdim) of the rotation part of a 4x4 affine."""
    rzs = np.asarray(affine, float)[:3, :3]
    zooms = np.sqrt((rzs * rzs).sum(0))
    zooms[zooms == 0] = 1.0
    r = rzs / zooms
    qfac = 1.0
    if np.linalg.det(r) < 0:
        r[:, 2] *= -1
        qfac = -1.0
    # nearest orthonormal matrix, then the standard matrix -> quaternion conversion (w >= 0)
    u, _, vt = np.linalg.svd(r)
    r = u @ vt
    k = np.array([[r[0, 0] - r[1, 1] - r[2, 2], 0, 0, 0], [r[0, 1] + r[1, 0], r[1, 1] - r[0, 0] - r[2, 2], 0, 0],
                  [r[0, 2] + r[2, 0], r[1, 2] + r[2, 1], r[2, 2] - r[0, 0] - r[1, 1], 0],
                  [r[2, 1] - r[1, 2], r[0, 2] - r[2, 0], r[1, 0] - r[0, 1], r[0, 0] + r[1, 1] + r[2, 2]]]) / 3.0
    vals, vecs = np.linalg.eigh(k + np.tril(k, -1).T)
    q = vecs[[3, 0, 1, 2], np.argmax(vals)]
    if q[0] < 0:
        q = -q
    return q[1:], qfac, zooms


def write_nifti(path, data, affine):
    """`nib.save(nib.Nifti1Image(data, affine), path)` without nibabel: single-file NIfTI-1 (`n+1`), data in Fortran order
    as the format prescribes, the affine in the sform (code 2 'aligned', what nibabel sets for an image built from an array
    and an affine) and its rotation in the qform parameters with code 0; gzip when the name ends in .gz."""
    data = np.asarray(data)
    if data.dtype not in _NIFTI_DT:
        raise ValueError(f"write_nifti: unsupported dtype {data.dtype}")
    if not 1 <= data.ndim <= 7:
        raise ValueError("write_nifti: 1 to 7 dimensions")
    affine = np.asarray(affine, float).reshape(4, 4)
    (qb, qc, qd), qfac, zooms = _quaternion_from_affine(affine)
    dim = [data.ndim] + list(data.shape) + [1] * (7 - data.ndim)
    pixdim = [qfac] + list(zooms) + [1.0] * 4
    hdr = bytearray(348)
    struct.pack_into("<i", hdr, 0, 348)
    struct.pack_into("<8h", hdr, 40, *dim)
    struct.pack_into("<h", hdr, 70, _NIFTI_DT[data.dtype])
    struct.pack_into("<h", hdr, 72, data.dtype.itemsize * 8)
    struct.pack_into("<8f", hdr, 76, *pixdim)
    struct.pack_into("<f", hdr, 108, 352.0)                 # vox_offset
    struct.pack_into("<f", hdr, 112, 1.0)                   # scl_slope
    struct.pack_into("<f", hdr, 116, 0.0)                   # scl_inter
    hdr[123] = 0                                            # xyzt_units: unknown, as a header built from scratch by nibabel
    struct.pack_into("<h", hdr, 252, 0)                     # qform_code: unknown
    struct.pack_into("<h", hdr, 254, 2)                     # sform_code: aligned
    struct.pack_into("<3f", hdr, 256, qb, qc, qd)
    struct.pack_into("<3f", hdr, 268, *affine[:3, 3])
    struct.pack_into("<4f", hdr, 280, *affine[0])
    struct.pack_into("<4f", hdr, 296, *affine[1])
    struct.pack_into("<4f", hdr, 312, *affine[2])
    hdr[344:348] = b"n+1\x00"
    payload = bytes(hdr) + b"\x00" * 4 + np.asfortranarray(data).tobytes(order="F")
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "wb") as f:
        f.write(payload)


def read_nifti(path):
    """Returns (data, affine) of a single-file NIfTI-1 image (sform when its code is set, else the qform, else pixdim)."""
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "rb") as f:
        blob = f.read()
    if struct.unpack_from("<i", blob, 0)[0] != 348:
        raise ValueError(f"{path}: not a little-endian NIfTI-1 file")
    dim = struct.unpack_from("<8h", blob, 40)
    dtype = _NIFTI_DT_INV[struct.unpack_from("<h", blob, 70)[0]]
    pixdim = struct.unpack_from("<8f", blob, 76)
    vox_offset = int(struct.unpack_from("<f", blob, 108)[0])
    slope, inter = struct.unpack_from("<2f", blob, 112)
    shape = tuple(dim[1:1 + dim[0]])
    data = np.frombuffer(blob, dtype=dtype, count=int(np.prod(shape)), offset=vox_offset).reshape(shape, order="F")
    if slope not in (0.0, 1.0) or inter != 0.0:
        data = data * slope + inter
    qform_code, sform_code = struct.unpack_from("<2h", blob, 252)
    affine = np.eye(4)
    if sform_code > 0:
        affine[0] = struct.unpack_from("<4f", blob, 280)
        affine[1] = struct.unpack_from("<4f", blob, 296)
        affine[2] = struct.unpack_from("<4f", blob, 312)
    elif qform_code > 0:
        b, c, d = struct.unpack_from("<3f", blob, 256)
        a = np.sqrt(max(0.0, 1.0 - (b * b + c * c + d * d)))
        r = np.array([[a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c)],
                      [2 * (b * c + a * d), a * a + c * c - b * b - d * d, 2 * (c * d - a * b)],
                      [2 * (b * d - a * c), 2 * (c * d + a * b), a * a + d * d - b * b - c * c]])
        qfac = -1.0 if pixdim[0] < 0 else 1.0
        affine[:3, :3] = r @ np.diag([pixdim[1], pixdim[2], pixdim[3] * qfac])
        affine[:3, 3] = struct.unpack_from("<3f", blob, 268)
    else:
        affine[:3, :3] = np.diag(pixdim[1:4])
    return np.ascontiguousarray(data), affine
