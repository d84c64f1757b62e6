"""Generates tests/golden/config1_*.npz and nnunet_ref.npz from the UNMODIFIED reference (/root/reference, CPU fp32).

BASELINE config 1 = a single 64^3 synthetic pair through convex_adam_pt (the reference's own test path,
tests/test_convex_adam_mind.py:22-42), with both parameter sets SURVEY.md 8(d) names:
    config1_ns       : grid_sp=2, disp_hw=6, 100 iterations (the north-star hyper-parameters)
    config1_defaults : the package defaults grid_sp=6, disp_hw=4, 80 iterations (convex_adam_MIND.py:64-82)
Stored per case (SURVEY 7.2 parity ladder, rung P3):
    field_n{1,15,N}  : the reference's field after 1, 15 and N iterations ([::2,::2,::2], float32)
    u_final          : disp_sample of the last forward (3,ha,wa,da) -- the field the objective is evaluated on
    objective        : data + regularisation loss of that forward (MIND.py:166-177), float64 re-evaluation
    envelope         : the reference's own sensitivity at N iterations: L_inf between the clean run and runs whose MIND
                       features were multiplied by (1 + 1.2e-7 * randn) (SURVEY App. B.4 protocol), one value per noise seed
    envelope_expf    : the same with amplitude 3e-7 = the measured difference between torch's SLEEF expf and libm / CUDA
                       expf on the MIND descriptors (the tolerance of the MIND parity tests): the perturbation ANY
                       independent implementation of utils.py:63 feeds the Adam loop
    objective_noisy  : the final objective of all those runs
nnunet_ref.npz: the reference's nnUNet `extract_features` (convex_adam_nnUNet.py:19-38, true half arithmetic on the
CPU, `.cuda()` shimmed to a no-op) on a seeded label pair, and the field of the reference's nnUNet `convex_adam`
(in-memory nibabel shim, `.half()` -> float because torch-CPU has no Half avg_pool3d: SURVEY 8(c)).

Only runnable in the build container (the reference cannot travel); the vectors it writes are committed.
    python tests/golden/make_golden_config1.py
"""
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.ref_shim import import_reference  # noqa: E402
from convexadam_b200.synth import make_label_pair, make_pair  # noqa: E402
from tests.objective import registration_objective  # noqa: E402

ref_utils, ref_mind = import_reference()
torch.set_num_threads(max(1, (os.cpu_count() or 2) // 2))   # the reference is bit-identical for 1 and N threads
CPU = torch.device("cpu")
HERE = os.path.dirname(os.path.abspath(__file__))


class Capture:
    """Records the Adam-grid features and the parameters of the last Adam step of one convex_adam_pt call."""

    def __enter__(self):
        self.pool = []
        self.last_p = None
        self._pool, self._adam = ref_mind.F.avg_pool3d, torch.optim.Adam
        C = self

        def avg_pool3d(x, *a, **k):
            out = C._pool(x, *a, **k)
            if x.dim() == 5 and x.shape[1] == 12:
                C.pool.append(out[0].detach().numpy().copy())
            return out

        class RecAdam(self._adam):
            def step(self_, closure=None):
                C.last_p = self_.param_groups[0]["params"][0].data[0].numpy().copy()
                return super().step(closure)

        ref_mind.F.avg_pool3d = avg_pool3d
        torch.optim.Adam = RecAdam
        return self

    def __exit__(self, *exc):
        ref_mind.F.avg_pool3d = self._pool
        torch.optim.Adam = self._adam


def box3_cubed(p):
    t = torch.from_numpy(p)[None]
    for _ in range(3):
        t = F.avg_pool3d(t, 3, stride=1, padding=1)
    return t[0].numpy()


def run_ref(fixed, moving, niter, kw, noise_seed=None, amplitude=1.2e-7):
    orig = ref_mind.MINDSSC
    if noise_seed is not None:
        g = torch.Generator().manual_seed(noise_seed)

        def noisy(img, r, d, device=None):
            out = orig(img, r, d, device=device)
            return out * (1 + amplitude * torch.randn(out.shape, generator=g))

        ref_mind.MINDSSC = noisy
    try:
        with Capture() as cap:
            field = ref_mind.convex_adam_pt(fixed, moving, selected_niter=niter, device=CPU, dtype=torch.float32, **kw)
    finally:
        ref_mind.MINDSSC = orig
    # avg_pool3d calls on 12-channel inputs: fix/mov at grid_sp, (mov/fix again for ic: no, correlate is called with the
    # same pooled tensors), then fix/mov at grid_sp_adam
    fix_a, mov_a = cap.pool[-2], cap.pool[-1]
    u = box3_cubed(cap.last_p)
    obj = registration_objective(u, fix_a, mov_a, float(kw["lambda_weight"]))
    return field, u, obj


def config1_case(name, seed, niter, kw, noise_seeds=(101, 102, 103, 104), expf_seeds=(201, 202, 203, 204)):
    shape = (64, 64, 64)
    fixed, moving = make_pair(shape, seed, disp_scale=0.08)
    out = dict(shape=np.array(shape), seed=np.array(seed), niter=np.array(niter), disp_scale=np.array(0.08),
               kw_keys=np.array(sorted(kw.keys())), kw_vals=np.array([float(kw[k]) for k in sorted(kw.keys())]))
    for n in (1, 15, niter):
        field, u, obj = run_ref(fixed, moving, n, kw)
        out[f"field_n{n}"] = field[::2, ::2, ::2].astype(np.float32)
        if n == niter:
            clean, out["u_final"], out["objective"] = field, u.astype(np.float32), np.array(obj)
    env, env2, objs = [], [], []
    for s in noise_seeds:
        field, _, obj = run_ref(fixed, moving, niter, kw, noise_seed=s)
        env.append(np.abs(field - clean).max())
        objs.append(obj)
    for s in expf_seeds:
        field, _, obj = run_ref(fixed, moving, niter, kw, noise_seed=s, amplitude=3e-7)
        env2.append(np.abs(field - clean).max())
        objs.append(obj)
    out["envelope"] = np.array(env)
    out["envelope_expf"] = np.array(env2)
    out["objective_noisy"] = np.array(objs)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(name, f"{os.path.getsize(path) / 1e6:.2f} MB; |field|max {np.abs(clean).max():.3f}; objective {float(out['objective']):.8f};",
          "envelope", env, env2, "noisy objectives rel", [abs(o / float(out["objective"]) - 1) for o in objs])


def nnunet_case():
    import importlib

    lf, lm = make_label_pair((36, 32, 40), n_labels=7, seed=5)
    lf[lf == 3] = 5          # label 3 absent from the fixed map only
    lm[lm == 2] = 0          # label 2 absent from the moving map only (the reference needs the largest label in both maps)
    lf[lf == 4] = 1
    lm[lm == 4] = 1          # label 4 absent from both: C = 6 of 7
    store = {}
    nib = sys.modules["nibabel"]
    files = {"fix.nii.gz": lf.numpy().astype(np.float64), "mov.nii.gz": lm.numpy().astype(np.float64)}

    class Img:
        def __init__(self, data, affine=None):
            self._d, self.affine = data, (np.eye(4) if affine is None else affine)

        def get_fdata(self):
            return self._d

    nib.load = lambda p: Img(files[os.path.basename(p)])
    nib.Nifti1Image = Img
    nib.save = lambda img, p: store.__setitem__("disp", img.get_fdata())
    saved = (torch.Tensor.cuda, torch.nn.Module.cuda, torch.cuda.synchronize, torch.Tensor.half)
    torch.Tensor.cuda = lambda s, *a, **k: s
    torch.nn.Module.cuda = lambda s, *a, **k: s
    torch.cuda.synchronize = lambda *a, **k: None
    try:
        ref_nn = importlib.import_module("convexAdam.convex_adam_nnUNet")
        ff, fm = ref_nn.extract_features(lf, lm)                      # true half arithmetic (CPU)
        assert ff.dtype == torch.float16
        torch.Tensor.half = lambda s: s.float()                       # fp32 oracle for the full path
        ref_nn.convex_adam("fix.nii.gz", "mov.nii.gz", 1.25, 2, 3, 6, 0, 2, True, "./")
        ff32, fm32 = ref_nn.extract_features(lf, lm)
    finally:
        torch.Tensor.cuda, torch.nn.Module.cuda, torch.cuda.synchronize, torch.Tensor.half = saved
    path = os.path.join(HERE, "nnunet_ref.npz")
    np.savez_compressed(path, labels_fix=lf.numpy().astype(np.int8), labels_mov=lm.numpy().astype(np.int8),
                        feat_fix_half=ff[0].numpy(), feat_mov_half=fm[0].numpy(),
                        feat_fix_f32_chanmax=ff32[0].reshape(ff32.shape[1], -1).max(1).values.numpy(),
                        params=np.array([1.25, 2, 3, 6, 0, 2]), field=store["disp"][::2, ::2, ::2].astype(np.float32))
    print("nnunet_ref", f"{os.path.getsize(path) / 1e6:.2f} MB; C = {ff.shape[1]}; |field|max {np.abs(store['disp']).max():.3f}")


if __name__ == "__main__":
    what = sys.argv[1:] or ["ns", "defaults", "nnunet"]
    if "ns" in what:
        config1_case("config1_ns", seed=21, niter=100,
                     kw=dict(mind_r=1, mind_d=2, lambda_weight=1.25, grid_sp=2, disp_hw=6, grid_sp_adam=2, ic=True, selected_smooth=0))
    if "defaults" in what:
        config1_case("config1_defaults", seed=22, niter=80,
                     kw=dict(mind_r=1, mind_d=2, lambda_weight=1.25, grid_sp=6, disp_hw=4, grid_sp_adam=2, ic=True, selected_smooth=0))
    if "nnunet" in what:
        nnunet_case()
