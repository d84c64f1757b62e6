"""Operator-level API, same names / arguments / shapes as the reference's
src/convexAdam/convex_adam_utils.py (MINDSSC :24, correlate :72, coupled_convex :93, inverse_consistency :114,
validate_image :268) -- each one a thin adapter onto a kernel of libconvexadam_b200.  CUDA tensors only.
"""
import warnings

import numpy as np
import torch

from ._native import get_lib

warnings.filterwarnings("ignore")  # the reference does this at import (utils.py:12)


def _f32(t):
    return t if t.dtype == torch.float32 else t.float()


def MINDSSC(img, radius=2, dilation=2, device="cuda"):
    """img (1,1,H,W,D) -> (1,12,H,W,D) float32 MIND-SSC descriptor (reference utils.py:24-68)."""
    lib = get_lib()
    if img.dim() != 5 or img.shape[0] != 1 or img.shape[1] != 1:
        raise ValueError("MINDSSC expects a (1,1,H,W,D) tensor")
    x = _f32(img[0, 0].to(device))
    out, _ = lib.mindssc_pooled(x, int(radius), int(dilation), 1)
    return out.unsqueeze(0)


def correlate(mind_fix, mind_mov, disp_hw, grid_sp, shape, ch=12):
    """SSD cost volume + argmin (reference utils.py:72-89).  mind_fix/mind_mov: (1,ch,h,w,d) with
    (h,w,d) = shape // grid_sp.  Returns ssd ((2*disp_hw+1)^3, h, w, d) and ssd_argmin (h,w,d) int64."""
    lib = get_lib()
    H, W, D = int(shape[0]), int(shape[1]), int(shape[2])
    h, w, d = H // grid_sp, W // grid_sp, D // grid_sp
    if tuple(mind_fix.shape) != (1, ch, h, w, d) or tuple(mind_mov.shape) != (1, ch, h, w, d):
        raise ValueError(f"correlate: expected features of shape {(1, ch, h, w, d)}, got {tuple(mind_fix.shape)} / {tuple(mind_mov.shape)}")
    dt = mind_fix.dtype
    ssd, am = lib.correlate(_f32(mind_fix[0]), _f32(mind_mov[0]), int(disp_hw))
    return (ssd if dt == torch.float32 else ssd.to(dt)), am


def coupled_convex(ssd, ssd_argmin, disp_mesh_t, grid_sp, shape):
    """Six coupled-convex passes (reference utils.py:93-109) -> disp_soft (1,3,h,w,d), coarse-voxel units."""
    lib = get_lib()
    L = ssd.shape[0]
    S = round(L ** (1.0 / 3.0))
    if S * S * S != L or S % 2 == 0:
        raise ValueError("coupled_convex: ssd.shape[0] must be (2*disp_hw+1)^3")
    dt = ssd.dtype
    ssd32 = _f32(ssd)
    # the kernels order (cost, label) pairs by the cost's bit pattern, which is the numerical order for non-negative floats
    # only: SSD costs are non-negative by construction, anything else (or a NaN) is rejected here instead of mis-sorted
    lo = ssd32.min()
    if not bool(lo >= 0):
        raise ValueError(f"coupled_convex: ssd must be non-negative (sum of squared differences); got min {float(lo)}")
    soft = lib.coupled_convex(ssd32, ssd_argmin, _f32(disp_mesh_t).reshape(3, -1), (S - 1) // 2)
    return (soft if dt == torch.float32 else soft.to(dt)).unsqueeze(0)


def inverse_consistency(disp_field1s, disp_field2s, iter=20):
    """Fixed-point inverse-consistency iterations (reference utils.py:114-129); fields (1,3,h,w,d), normalised,
    channel 0 = x (last array axis)."""
    lib = get_lib()
    dt = disp_field1s.dtype
    o1, o2 = lib.inverse_consistency(_f32(disp_field1s[0]), _f32(disp_field2s[0]), int(iter))
    if dt != torch.float32:
        o1, o2 = o1.to(dt), o2.to(dt)
    return o1.unsqueeze(0), o2.unsqueeze(0)


def validate_image(img, dtype=float) -> torch.Tensor:
    """Same acceptance rules and error as the reference (utils.py:268-279); SimpleITK / nibabel are optional."""
    if not isinstance(img, torch.Tensor):
        try:
            import SimpleITK as sitk
            if isinstance(img, sitk.Image):
                img = sitk.GetArrayFromImage(img)
        except ImportError:
            pass
        try:
            import nibabel as nib
            if isinstance(img, nib.Nifti1Image):
                img = img.get_fdata()
        except ImportError:
            pass
        if isinstance(img, np.ndarray):
            img = torch.from_numpy(img.astype(dtype))
        else:
            raise ValueError("Input image must be a torch.Tensor, a numpy.ndarray or a SimpleITK.Image")
    return img
