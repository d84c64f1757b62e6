"""Array-level core of the reference's src/convexAdam/convex_adam_translation.py (translation-only registration).

The reference wrapper (:57-114) resamples two SimpleITK images to 1 mm, runs `convex_adam_pt` with its defaults, averages
the dense field (optionally inside a segmentation) and turns the mean into a whole-voxel translation of the moving
image's origin.  The SimpleITK geometry (resample_img, resample_moving_to_fixed, origin edit) is CPU work outside the hot
path and is not rebuilt; what IS on the path -- registration + mean -- runs here on the device, and
the 189 MB dense field never leaves it: `ca_field_mean` reduces it to three numbers.
"""
from typing import Iterable, Optional

import numpy as np
import torch

from . import pipeline
from ._native import get_lib
from .convex_adam_utils import validate_image


def index_translation_to_world_translation(index_translation: Iterable[float], direction: Iterable[float]) -> np.ndarray:
    """A translation along the image grid (mm) expressed in world coordinates: direction cosines times the vector
    (same contract as the reference helper of that name, :12-29)."""
    cosines = np.asarray(list(direction), dtype=np.float64)
    dim = int(round(np.sqrt(cosines.size)))
    return cosines.reshape(dim, dim).dot(np.asarray(list(index_translation), dtype=np.float64))


def field_to_translation(translation_zyx, spacing_zyx) -> tuple:
    """Mean displacement on the 1 mm grid (array-axis order, mm) -> the shift the reference applies to the moving image
    (:99-104): snapped to whole voxels of the moving image's own spacing, given back in mm as (x, y, z)."""
    sp = np.asarray(spacing_zyx, dtype=np.float64)
    whole_voxels = np.rint(np.asarray(translation_zyx, dtype=np.float64) / sp)       # np.round(., 0): half to even, like rint
    return tuple((whole_voxels * sp)[::-1].tolist())


def mean_displacement(img_fixed, img_moving, segmentation=None, device=None, dtype=torch.float16, **kw) -> np.ndarray:
    """np.mean(convex_adam_pt(img_fixed, img_moving, **kw)[seg > 0], axis=0) (reference :83-96) without moving the dense
    field to the host.  img_*: (H,W,D) volumes already on the common 1 mm grid; segmentation: optional (H,W,D) array,
    voxels > 0 count.  kw: convex_adam_pt's hyper-parameters (defaults = the reference's).  Returns 3 float64 values
    (array-axis order)."""
    lib = get_lib()
    dev = torch.device(device if device is not None else "cuda")
    fixed = validate_image(img_fixed).float()
    moving = validate_image(img_moving).float()
    names = ("mind_r", "mind_d", "lambda_weight", "grid_sp", "disp_hw", "selected_niter", "selected_smooth", "grid_sp_adam", "ic")
    unknown = set(kw) - set(names)
    if unknown:
        raise TypeError(f"unexpected arguments {sorted(unknown)}")
    prm = pipeline.Params(**{k: kw[k] for k in names if k in kw})
    with torch.cuda.device(dev):
        disp = pipeline.register_mind(fixed.to(dev).contiguous(), moving.to(dev).contiguous(), prm, lib)
        mask = None
        if segmentation is not None:
            seg = validate_image(segmentation)
            if tuple(seg.shape) != tuple(disp.shape[1:]):
                raise ValueError("segmentation must have the shape of the images")
            mask = (seg > 0).to(torch.uint8).to(dev).contiguous()
        out = lib.field_mean(disp, mask, round_half=(dtype == torch.float16))
        return out[:3].cpu().numpy()


def convex_adam_translation_arrays(img_fixed, img_moving, moving_spacing_zyx, segmentation: Optional[np.ndarray] = None,
                                   device=None, **kw) -> tuple:
    """(x, y, z) translation in mm as `convex_adam_translation` computes it (:83-104), from arrays on the common 1 mm
    grid and the moving image's spacing (array-axis order)."""
    return field_to_translation(mean_displacement(img_fixed, img_moving, segmentation, device=device, **kw), moving_spacing_zyx)
