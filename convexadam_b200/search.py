"""Convex-stage sweeps over hyper-parameter settings with the features computed once.

The reference's self-configuring search (self_configuring/convex_run_withconfig.py:78-156) re-runs MIND, correlate and
coupled_convex for every (grid_sp, disp_hw) setting of every pair.  MIND does not depend on the setting: here the
full-resolution descriptors are computed once per image, pooled once per distinct grid_sp, and everything stays on the
device; each returned field is bit-identical to `convex_adam_pt(..., grid_sp=gs, disp_hw=hw, lambda_weight=0)`.
"""
import numpy as np
import torch

from . import pipeline
from ._native import get_lib


def convex_fields(img_fixed, img_moving, settings, mind_r=1, mind_d=2, ic=True, device=None, to_host=True):
    """settings: iterable of (grid_sp, disp_hw).  Returns one field per setting, in order: (H,W,D,3) float64 numpy arrays
    (to_host=True) or (3,H,W,D) float32 CUDA tensors.  With ic=False the reference returns the coarse, un-upsampled
    field (MIND.py:143-144) and so does this."""
    lib = get_lib()
    dev = torch.device(device if device is not None else "cuda")
    fixed = torch.as_tensor(img_fixed).float()
    moving = torch.as_tensor(img_moving).float()
    if fixed.dim() != 3 or fixed.shape != moving.shape:
        raise ValueError("img_fixed and img_moving must be 3-D volumes of the same shape")
    H, W, D = fixed.shape
    settings = [(int(gs), int(hw)) for gs, hw in settings]
    out = []
    with torch.cuda.device(dev):
        full_f, _ = lib.mindssc_pooled(fixed.to(dev).contiguous(), mind_r, mind_d, 1)
        full_m, _ = lib.mindssc_pooled(moving.to(dev).contiguous(), mind_r, mind_d, 1)
        pooled = {}
        ssd_buf = None
        for gs, hw in settings:
            if gs not in pooled:
                pooled[gs] = (full_f if gs == 1 else lib.avgpool3d(full_f, gs, gs), full_m if gs == 1 else lib.avgpool3d(full_m, gs, gs))
            fs, ms = pooled[gs]
            prm = pipeline.Params(mind_r, mind_d, 0.0, gs, hw, 1, 0, 2, ic)
            L = (2 * hw + 1) ** 3
            if ssd_buf is None or ssd_buf.numel() < L * fs[0].numel():
                ssd_buf = None                       # release before the larger volume is allocated
                ssd_buf = torch.empty(L * fs[0].numel(), dtype=torch.float32, device=dev)
            vol = ssd_buf[:L * fs[0].numel()].view((L,) + tuple(fs.shape[1:]))
            disp = pipeline.convex_stage(lib, fs, ms, (H, W, D), prm, vol)
            if to_host:
                out.append(lib.pack_field(disp).cpu().numpy())
            else:
                out.append(disp)
    return out
