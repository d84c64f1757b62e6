"""Drop-in for the reference's src/convexAdam/convex_adam_nnUNet.py: extract_features (:19-38) and convex_adam
(:41-159, file-in / disp.nii.gz-out), plus an in-memory `convex_adam_pt` the reference lacks.  Same pipeline as
the MIND path with C = number of labels present in either prediction; the literal `*12` of the loss is kept
whatever C is (nnUNet.py:127).
"""
import argparse
import os
import warnings

import numpy as np
import torch

from . import pipeline
from ._native import get_lib

warnings.filterwarnings("ignore")


def _label_tables(pred_fixed, pred_moving):
    """Host-side bookkeeping of nnUNet.py:22-33: labels present in either map (ascending), class weights
    (1/(count+eps))^0.3 normalised by their mean, all in torch fp32 exactly as the reference computes them."""
    eps = 1e-32
    pf = pred_fixed.long().reshape(-1).cpu()
    pm = pred_moving.long().reshape(-1).cpu()
    nb = int(max(pf.max(), pm.max())) + 1
    combined = torch.bincount(pf, minlength=nb) + torch.bincount(pm, minlength=nb)
    pos = torch.nonzero(combined).reshape(-1)
    weight = 1 / (combined[pos] + eps).float().pow(0.3)
    weight /= weight.mean()
    lut = torch.full((nb,), -1, dtype=torch.int32)
    lut[pos] = torch.arange(len(pos), dtype=torch.int32)
    return pos, weight, lut


def half_weights(weight):
    """The value a present class takes in the reference's features, `10 * (onehot.float() * weight).half()` (nnUNet.py:35-36):
    the weight rounded to half, then multiplied by 10 in half arithmetic.  Returned as float32 (exactly representable)."""
    return (10 * weight.half()).half().float()


def extract_features(pred_fixed, pred_moving):
    """Weighted one-hot features (1,C,H,W,D) of both label maps in half precision (reference nnUNet.py:19-38).  Like the
    reference, any leading singleton dimensions are accepted (`shape[-3:]`, `view(1,H,W,D)`)."""
    lib = get_lib()
    dev = torch.device("cuda")
    pred_fixed = torch.as_tensor(pred_fixed)
    pred_moving = torch.as_tensor(pred_moving)
    H, W, D = pred_fixed.shape[-3:]
    pos, weight, lut = _label_tables(pred_fixed, pred_moving)
    w = half_weights(weight)
    ff = lib.onehot_pooled(pred_fixed.float().reshape(H, W, D).to(dev).contiguous(), lut.to(dev), w.to(dev), 1)
    fm = lib.onehot_pooled(pred_moving.float().reshape(H, W, D).to(dev).contiguous(), lut.to(dev), w.to(dev), 1)
    return ff.half().unsqueeze(0), fm.half().unsqueeze(0)


def convex_adam_pt(pred_fixed, pred_moving, lambda_weight=1.25, grid_sp=6, disp_hw=4, selected_niter=80,
                   selected_smooth=0, grid_sp_adam=2, ic=True, dtype=torch.float16, device=None):
    """In-memory form of convex_adam below: label maps (H,W,D) -> (H,W,D,3) float64 displacement field."""
    lib = get_lib()
    dev = torch.device(device if device is not None else "cuda")
    if dev.type != "cuda":
        raise RuntimeError("convexadam_b200 runs on CUDA devices only")
    pred_fixed = torch.as_tensor(pred_fixed).float()
    pred_moving = torch.as_tensor(pred_moving).float()
    H, W, D = pred_fixed.shape[-3:]
    pos, weight, lut = _label_tables(pred_fixed, pred_moving)
    if dtype == torch.float16:
        w = half_weights(weight)
    else:
        w = 10 * weight
    # the reference honours selected_smooth in {3,5} only (nnUNet.py:136-144)
    smooth = selected_smooth if selected_smooth in (3, 5) else 0
    # one-hot label features give many exactly equal costs: the certified mode would re-decide them one candidate at a
    # time (same selections, no speed-up there), so this path takes the exact cost volume by default
    prm = pipeline.Params(1, 2, lambda_weight, grid_sp, disp_hw, selected_niter, smooth, grid_sp_adam, ic,
                          mode=os.environ.get("CONVEXADAM_B200_MODE", "exact"))
    with torch.cuda.device(dev):
        lf = pred_fixed.reshape(H, W, D).to(dev).contiguous()
        lm = pred_moving.reshape(H, W, D).to(dev).contiguous()
        lut_d, w_d = lut.to(dev), w.to(dev)
        fix_s = lib.onehot_pooled(lf, lut_d, w_d, grid_sp)
        mov_s = lib.onehot_pooled(lm, lut_d, w_d, grid_sp)
        if grid_sp_adam == grid_sp:
            fix_a, mov_a = fix_s, mov_s
        else:
            fix_a = lib.onehot_pooled(lf, lut_d, w_d, grid_sp_adam)
            mov_a = lib.onehot_pooled(lm, lut_d, w_d, grid_sp_adam)
        disp = pipeline.register_features(lib, fix_s, mov_s, fix_a, mov_a, (H, W, D), prm)
        packed = lib.pack_field(disp, round_half=(dtype == torch.float16))
        host = torch.empty(packed.shape, dtype=torch.float64, pin_memory=True)
        host.copy_(packed, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    return host.numpy()


def convex_adam(path_pred_fixed, path_pred_moving, lambda_weight, grid_sp, disp_hw, selected_niter, selected_smooth,
                grid_sp_adam=2, ic=True, result_path="./"):
    """Reference signature (nnUNet.py:41-50): reads two label maps, writes result_path/disp.nii.gz (nibabel when installed,
    else this package's own NIfTI-1 reader / writer)."""
    from .convex_adam_MIND import _nifti_io
    load, save = _nifti_io()
    fixed_data, affine = load(path_pred_fixed)
    pred_fixed = torch.from_numpy(np.asarray(fixed_data)).float()
    pred_moving = torch.from_numpy(np.asarray(load(path_pred_moving)[0])).float()
    displacements = convex_adam_pt(pred_fixed, pred_moving, lambda_weight, grid_sp, disp_hw, selected_niter,
                                   selected_smooth, grid_sp_adam, ic)
    save(os.path.join(result_path, "disp.nii.gz"), displacements, affine)


def _main():
    parser = argparse.ArgumentParser()
    parser.add_argument("-f", "--path_pred_fixed", type=str, required=True)
    parser.add_argument("-m", "--path_pred_moving", type=str, required=True)
    parser.add_argument("--lambda_weight", type=float, default=1.25)
    parser.add_argument("--grid_sp", type=int, default=6)
    parser.add_argument("--disp_hw", type=int, default=4)
    parser.add_argument("--selected_niter", type=int, default=80)
    parser.add_argument("--selected_smooth", type=int, default=0)
    parser.add_argument("--grid_sp_adam", type=int, default=2)
    parser.add_argument("--ic", choices=("True", "False"), default="True")
    parser.add_argument("--result_path", type=str, default="./")
    a = parser.parse_args()
    convex_adam(a.path_pred_fixed, a.path_pred_moving, a.lambda_weight, a.grid_sp, a.disp_hw, a.selected_niter,
                a.selected_smooth, a.grid_sp_adam, a.ic == "True", a.result_path)


if __name__ == "__main__":
    _main()
