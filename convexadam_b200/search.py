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


def kovesi_widths(sigma, n=4):
    """Widths of the n average pools whose sequence approximates a Gaussian of standard deviation sigma (Kovesi's "fast almost
    Gaussian filtering"): m pools of the lower odd width w_l (none when w_l = 1), then n - m of the upper width w_u --
    the list `kovesi_spline(sigma, n)` builds at convexAdam_hyper_util.py:475-488."""
    w_ideal = np.sqrt(12 * sigma ** 2 / n + 1)
    w_u = int(np.ceil((w_ideal - 1) / 2) * 2 + 1)
    w_l = max(w_u - 2, 1)
    m = int(np.round((12 * sigma ** 2 - n * w_l ** 2 - 4 * n * w_l - 3 * n) / (-4 * w_l - 4)))
    return ([w_l] * m if w_l > 1 else []) + [w_u] * (n - m)


def gaussian_weights(sigma):
    """The normalised filter weights of the reference's GaussianSmoothing(sigma), computed with the same float32 torch operations
    (convexAdam_hyper_util.py:458-464): N = 2 ceil(1.5 sigma) + 1 taps."""
    sg = torch.tensor([float(sigma)])
    N = int(torch.ceil(sg * 3.0 / 2.0).long().item()) * 2 + 1
    weight = torch.exp(-torch.pow(torch.linspace(-(N // 2), N // 2, N), 2) / (2 * torch.pow(sg, 2)))
    weight /= weight.sum()
    return [float(x) for x in weight]


def adam_fields(img_fixed, img_moving, adam_settings, snapshots=(80,), grid_sp=6, disp_hw=4, selected_smooth=0, mind_r=1, mind_d=2,
                ic=True, device=None, to_host=True, smoother=None):
    """The Adam stage for a list of settings on top of ONE convex stage (second half).

    The reference's Adam search (self_configuring/adam_run_withconfig_shiftSpline.py:159-264) restarts from the stored
    convex field for every (grid_sp_adam, ..., lambda_weight) setting, recomputes the features each time and evaluates the
    field at several iteration counts of one run (iterations 60, 80, 100, 120 there).  Here MIND runs once per image, the
    convex stage once, the features are pooled once per distinct grid_sp_adam, and each run is continued from snapshot to
    snapshot (`ca_adam_run` is resumable: the state (p, m, v) and the step index carry over).

    adam_settings: iterable of (grid_sp_adam, lambda_weight); snapshots: increasing iteration counts.
    Returns {(grid_sp_adam, lambda_weight, niter): field}; every field is bit-identical to
    `convex_adam_pt(..., grid_sp_adam=gsa, lambda_weight=lam, selected_niter=niter)`.
    smoother: None = the reference's B^3 (MIND.py:166); ("kovesi", sigma[, n]) or a sequence of odd pool widths = the
    pool-sequence smoother of the search (`avgs[avg_n]`, adam_run_withconfig_shiftSpline.py:140-141,217; `kovesi_widths`);
    ("gauss", sigma) = its GaussianSmoothing(sigma) (replicate-padded separable filter, hyper_util.py:454-473)."""
    lib = get_lib()
    dev = torch.device(device if device is not None else "cuda")
    fixed = torch.as_tensor(img_fixed).float()
    moving = torch.as_tensor(img_moving).float()
    if fixed.dim() != 3 or fixed.shape != moving.shape:
        raise ValueError("img_fixed and img_moving must be 3-D volumes of the same shape")
    pools = gauss = None
    if smoother is not None and len(smoother) == 2 and smoother[0] == "gauss":
        gauss = gaussian_weights(smoother[1])
    elif smoother is not None:
        pools = kovesi_widths(*smoother[1:]) if (len(smoother) >= 2 and smoother[0] == "kovesi") else [int(w) for w in smoother]
        if not pools or any(w < 1 or w % 2 == 0 for w in pools):
            raise ValueError("adam_fields: the smoother must be a non-empty sequence of odd pool widths")
    snapshots = [int(n) for n in snapshots]
    if not snapshots or snapshots != sorted(set(snapshots)) or snapshots[0] < 1:
        raise ValueError("snapshots must be increasing iteration counts >= 1")
    H, W, D = fixed.shape
    out = {}
    with torch.cuda.device(dev):
        full_f, _ = lib.mindssc_pooled(fixed.to(dev).contiguous(), mind_r, mind_d, 1)
        full_m, _ = lib.mindssc_pooled(moving.to(dev).contiguous(), mind_r, mind_d, 1)
        pooled = {}

        def feats(gs):
            if gs not in pooled:
                pooled[gs] = (full_f if gs == 1 else lib.avgpool3d(full_f, gs, gs), full_m if gs == 1 else lib.avgpool3d(full_m, gs, gs))
            return pooled[gs]

        fs, ms = feats(int(grid_sp))
        base = pipeline.Params(mind_r, mind_d, 1.0, int(grid_sp), int(disp_hw), snapshots[-1], int(selected_smooth), 2, ic)
        disp0 = pipeline.convex_stage(lib, fs, ms, (H, W, D), base)
        for gsa, lam in adam_settings:
            gsa, lam = int(gsa), float(lam)
            if lam <= 0:
                raise ValueError("adam_fields: lambda_weight must be > 0 (lambda_weight = 0 skips the Adam stage: use convex_fields)")
            fa, ma = feats(gsa)
            prm = pipeline.Params(mind_r, mind_d, lam, int(grid_sp), int(disp_hw), snapshots[-1], int(selected_smooth), gsa, ic)
            p = pipeline.adam_start(lib, disp0, fa.shape[1:], prm)
            m, v = torch.zeros_like(p), torch.zeros_like(p)
            done = 0
            for n in snapshots:
                u = lib.adam_run(fa, ma, lam, p, m, v, done + 1, n - done, pools=pools, gauss=gauss)
                done = n
                disp = pipeline.adam_finish(lib, u, (H, W, D), prm)
                out[(gsa, lam, n)] = lib.pack_field(disp).cpu().numpy() if to_host else disp
    return out
