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
    """Translation along the image grid (mm) -> world coordinates (reference :12-29)."""
    dimension = int(np.sqrt(len(direction)))
    direction_matrix = np.array(direction).reshape((dimension, dimension))
    return direction_matrix @ np.array(index_translation)


def field_to_translation(translation_zyx, spacing_zyx) -> tuple:
    """Mean displacement (voxels of the 1 mm grid = mm, array-axis order) -> the translation the reference applies
    (:99-104): rounded to whole voxels of the moving image, in mm, reversed to (x, y, z)."""
    spacing_zyx = np.asarray(spacing_zyx, dtype=np.float64)
    translation_ijk = np.asarray(translation_zyx, dtype=np.float64) / spacing_zyx
    translation_ijk_voxels = np.round(translation_ijk, decimals=0)
    translation_ijk_mm = translation_ijk_voxels * spacing_zyx
    return tuple(list(translation_ijk_mm[::-1]))


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
