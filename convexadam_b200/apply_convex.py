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
    warped image as a numpy array of moving's floating dtype.  warped(x) = moving(x + disp(x)), linear, 0 outside."""
    lib = get_lib()
    dev = torch.device(device if device is not None else "cuda")
    moving = validate_image(moving)
    disp = validate_image(disp)
    if moving.dtype not in (torch.float32, torch.float64):
        moving = moving.double()                      # scipy interpolates integer images in float64 as well
    with torch.cuda.device(dev):
        out = lib.apply_convex(disp.to(dev, torch.float64), moving.to(dev))
        return out.cpu().numpy()
