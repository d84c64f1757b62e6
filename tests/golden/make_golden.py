"""Generates tests/golden/*.npz by running the UNMODIFIED reference (/root/reference, CPU fp32) on seeded
synthetic pairs, capturing every stage boundary of convex_adam_pt (SURVEY.md §7.2 parity ladder).

Only runnable in the build container (the reference cannot travel); the vectors it writes are committed.
    python tests/golden/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.ref_shim import import_reference  # noqa: E402
from convexadam_b200.synth import make_pair  # noqa: E402

ref_utils, ref_mind = import_reference()
torch.set_num_threads(1)   # single thread: the torch `sum(0)` tail (n % 32 columns) is then the only order quirk
CPU = torch.device("cpu")
HERE = os.path.dirname(os.path.abspath(__file__))


class Recorder:
    def __init__(self):
        self.rec = {}
        self.adam = []

    def patch(self):
        R = self
        self._orig = dict(correlate=ref_mind.correlate, coupled=ref_mind.coupled_convex, ic=ref_mind.inverse_consistency,
                          mind=ref_mind.MINDSSC, adam=torch.optim.Adam, interp=F.interpolate)
        n = {"corr": 0, "cc": 0, "mind": 0, "interp": 0}

        def correlate(*a, **k):
            ssd, am = R._orig["correlate"](*a, **k)
            sfx = "" if n["corr"] == 0 else "_"
            R.rec["fix_s" if n["corr"] == 0 else "mov_s_check"] = a[0][0].numpy().copy()
            if n["corr"] == 0:
                R.rec["mov_s"] = a[1][0].numpy().copy()
            R.rec["ssd" + sfx] = ssd.numpy().copy()
            R.rec["argmin" + sfx] = am.numpy().copy()
            n["corr"] += 1
            return ssd, am

        def coupled(ssd, am, mesh, gs, shape):
            sfx = "" if n["cc"] == 0 else "_"
            R.rec["mesh"] = mesh.numpy().reshape(3, -1).copy()
            out = R._orig["coupled"](ssd, am, mesh, gs, shape)
            R.rec["soft" + sfx] = out[0].numpy().copy()
            n["cc"] += 1
            return out

        def ic(a, b, iter=20):
            o1, o2 = R._orig["ic"](a, b, iter=iter)
            R.rec["ice"] = o1[0].numpy().copy()
            return o1, o2

        def mind(img, r, d, device=None):
            out = R._orig["mind"](img, r, d, device=device)
            R.rec["mind_fix" if n["mind"] == 0 else "mind_mov"] = out[0].numpy().copy()
            n["mind"] += 1
            return out

        def interp(x, *a, **k):
            out = R._orig["interp"](x, *a, **k)
            R.rec[f"interp{n['interp']}"] = out[0].detach().numpy().copy()
            n["interp"] += 1
            return out

        class RecAdam(R._orig["adam"]):
            def step(self_, closure=None):
                p = self_.param_groups[0]["params"][0]
                st = self_.state.get(p, {})
                R.adam.append(dict(p=p.data[0].numpy().copy(), g=p.grad[0].numpy().copy(),
                                   m=st["exp_avg"][0].numpy().copy() if st else np.zeros_like(p.data[0].numpy()),
                                   v=st["exp_avg_sq"][0].numpy().copy() if st else np.zeros_like(p.data[0].numpy())))
                r = super().step(closure)
                R.adam[-1]["p_next"] = p.data[0].numpy().copy()
                return r

        ref_mind.correlate = correlate
        ref_mind.coupled_convex = coupled
        ref_mind.inverse_consistency = ic
        ref_mind.MINDSSC = mind
        ref_mind.F.interpolate = interp
        torch.optim.Adam = RecAdam

    def unpatch(self):
        ref_mind.correlate = self._orig["correlate"]
        ref_mind.coupled_convex = self._orig["coupled"]
        ref_mind.inverse_consistency = self._orig["ic"]
        ref_mind.MINDSSC = self._orig["mind"]
        ref_mind.F.interpolate = self._orig["interp"]
        torch.optim.Adam = self._orig["adam"]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def sample_entries(a, k, seed):
    rs = np.random.RandomState(seed)
    idx = rs.randint(0, a.size, size=k).astype(np.int64)
    return idx, a.reshape(-1)[idx].copy()


def run_case(name, shape, seed, niter, keep_iters, disp_scale=0.04, **kw):
    fixed, moving = make_pair(shape, seed, disp_scale=disp_scale)
    R = Recorder()
    R.patch()
    try:
        field = ref_mind.convex_adam_pt(fixed, moving, selected_niter=niter, device=CPU, dtype=torch.float32, **kw)
        # second run with selected_niter=1: the convex stage + B^3 only
    finally:
        R.unpatch()
    field1 = ref_mind.convex_adam_pt(fixed, moving, selected_niter=1, device=CPU, dtype=torch.float32, **kw)
    rec = R.rec
    out = dict(
        shape=np.array(shape), seed=np.array(seed), niter=np.array(niter),
        kw_keys=np.array(sorted(kw.keys())), kw_vals=np.array([float(kw[k]) for k in sorted(kw.keys())]),
        fixed_sha=np.array(sha(fixed.numpy())), moving_sha=np.array(sha(moving.numpy())),
        fixed=fixed.numpy(), moving=moving.numpy(), disp_scale=np.array(disp_scale),
        fix_s=rec["fix_s"], mov_s=rec["mov_s"], mesh=rec["mesh"],
        ssd_sha=np.array(sha(rec["ssd"])), ssd__sha=np.array(sha(rec["ssd_"])),
        argmin=rec["argmin"].astype(np.int16), argmin_=rec["argmin_"].astype(np.int16),
        soft=rec["soft"], soft_=rec["soft_"], ice=rec["ice"],
        field=field[::2, ::2, ::2].astype(np.float32), field_sha=np.array(sha(field.astype(np.float32))),
        field_niter1=field1[::2, ::2, ::2].astype(np.float32), field_niter1_sha=np.array(sha(field1.astype(np.float32))),
        field_absmax=np.array(np.abs(field).max()),
    )
    i, v = sample_entries(rec["ssd"], 8192, 1); out["ssd_idx"] = i; out["ssd_val"] = v
    i, v = sample_entries(rec["ssd_"], 8192, 2); out["ssd__idx"] = i; out["ssd__val"] = v
    # MIND full-res features: sampled entries (tolerance check only: libm/SLEEF expf differ)
    i, v = sample_entries(rec["mind_fix"], 16384, 3); out["mind_fix_idx"] = i; out["mind_fix_val"] = v
    i, v = sample_entries(rec["mind_mov"], 16384, 4); out["mind_mov_idx"] = i; out["mind_mov_val"] = v
    out["disp_hr0"] = rec["interp0"][:, ::2, ::2, ::2]           # upsampled IC field (MIND.py:141)
    out["p0_src"] = rec["interp1"]                                # disp_lr (MIND.py:153), p0 = this / grid_sp_adam
    for k in keep_iters:
        for key in ("p", "g", "m", "v", "p_next"):
            out[f"adam{k}_{key}"] = R.adam[k][key]
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(name, "written", os.path.getsize(path) / 1e6, "MB; |field|max", float(np.abs(field).max()),
          "mean", field.mean((0, 1, 2)))


if __name__ == "__main__":
    # A: the north-star hyper-parameters on a ragged small volume (coarse grid 20x18x22, L = 343)
    run_case("pair_gs2_hw3", (40, 36, 44), seed=11, niter=12, keep_iters=(0, 5, 11), disp_scale=0.15,
             mind_r=1, mind_d=2, lambda_weight=1.25, grid_sp=2, disp_hw=3, grid_sp_adam=2, ic=True, selected_smooth=0)
    # B: package defaults (grid_sp=6, disp_hw=4, grid_sp_adam=2) incl. the odd smoothing kernel
    run_case("pair_defaults", (50, 44, 56), seed=12, niter=8, keep_iters=(0, 7), disp_scale=0.3,
             mind_r=1, mind_d=2, lambda_weight=1.25, grid_sp=6, disp_hw=4, grid_sp_adam=2, ic=True, selected_smooth=3)
