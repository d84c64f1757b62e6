"""Drop-in for the reference's src/convexAdam/apply_convex.py: `apply_convex(disp, moving)` (:13-24), the warp of the
moving image by the field convex_adam_pt returns, on the GPU (ca_apply_convex) instead of scipy.ndimage.map_coordinates.
`apply_convex_original_moving` (:27-78) is SimpleITK resampling of the field into another image's physical space and
stays out of scope (CPU geometry)."""
import numpy as np
import torch

from ._native import get_lib
from .convex_adam_utils import validate_image


def apply_convex(disp, moving, device=None) -> np.ndarray:
    """Same arguments and result as the reference: disp (H,W,D,3), moving (H,W,D), both torch / numpy / sitk; returns the
    warped image as a numpy array of moving's dtype (integer images: interpolated in float64, rounded back, as scipy does).  warped(x) = moving(x + disp(x)), linear, 0 outside."""
    lib = get_lib()
    dev = torch.device(device if device is not None else "cuda")
    # scipy.ndimage.map_coordinates returns the INPUT's dtype: an integer image is interpolated in float64 and rounded back
    in_dtype = getattr(moving, "dtype", None)
    int_dtype = None
    if isinstance(in_dtype, np.dtype) and in_dtype.kind in "iu":
        int_dtype = in_dtype
    elif isinstance(in_dtype, torch.dtype) and not in_dtype.is_floating_point and in_dtype != torch.bool:
        int_dtype = np.dtype(str(in_dtype).replace("torch.", ""))
    moving = validate_image(moving)
    disp = validate_image(disp)
    if int_dtype is not None or moving.dtype not in (torch.float32, torch.float64):
        moving = moving.double()
    with torch.cuda.device(dev):
        out = lib.apply_convex(disp.to(dev, torch.float64), moving.to(dev)).cpu().numpy()
    if int_dtype is not None:
        info = np.iinfo(int_dtype)
        out = np.clip(np.rint(out), info.min, info.max).astype(int_dtype)     # scipy: round half to even, saturate
    return out
