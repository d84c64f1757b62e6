"""Config 4 (nnUNet one-hot features, 16 labels, 192^3, disp_hw=6, grid_sp=2): convex stage in both modes + flagged counts."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from convexadam_b200._native import get_lib  # noqa: E402
from convexadam_b200.convex_adam_nnUNet import _label_tables, convex_adam_pt  # noqa: E402
from convexadam_b200.synth import make_label_pair  # noqa: E402

lib = get_lib()
dev = torch.device("cuda:0")
shape = (192, 192, 192)
lf, lm = make_label_pair(shape, n_labels=16, seed=1)
pos, weight, lut = _label_tables(lf, lm)
w = (10 * weight).to(dev)
fs = lib.onehot_pooled(lf.to(dev), lut.to(dev), w, 2)
ms = lib.onehot_pooled(lm.to(dev), lut.to(dev), w, 2)
mesh = lib.disp_mesh(6, dev)
print("features", tuple(fs.shape))


def timeit(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        r = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, r


ssd = torch.empty((2197,) + tuple(fs.shape[1:]), dtype=torch.float32, device=dev)
t_e, _ = timeit(lambda: lib.correlate(fs, ms, 6, out=ssd, want_argmin=False))
t_ce, (soft_e, sel_e) = timeit(lambda: lib.coupled_convex(ssd, None, mesh, 6, want_sel=True))
am = torch.argmin(ssd[:, :48], 0)
t_f, _ = timeit(lambda: lib.correlate(fs, ms, 6, out=ssd, fast=True))
t_cf, (soft_c, sel_c, fl) = timeit(lambda: lib.coupled_convex_certified(ssd, fs, ms, mesh, 6, want_sel=True, want_flagged=True))
print(f"exact: correlate {t_e:.2f} ms + coupled {t_ce:.2f} ms | certified: correlate {t_f:.2f} ms + coupled {t_cf:.2f} ms")
print("flagged per selection:", fl.tolist(), "of", soft_e[0].numel())
print("selections equal:", bool((sel_c[1:] == sel_e).all()), "soft bits equal:", bool((soft_c.view(torch.int32) == soft_e.view(torch.int32)).all()))
for mode in ("certified", "exact"):
    os.environ["CONVEXADAM_B200_MODE"] = mode
    convex_adam_pt(lf, lm, 1.25, 2, 6, 100, 0, 2, True, dtype=torch.float32)
    t0 = time.perf_counter(); out = convex_adam_pt(lf, lm, 1.25, 2, 6, 100, 0, 2, True, dtype=torch.float32); t1 = time.perf_counter()
    print(f"nnUNet convex_adam_pt ({mode}), 192^3, 16 labels, hw=6 gs=2 100 it.: {1e3 * (t1 - t0):.1f} ms end to end; field range {out.min():.2f}..{out.max():.2f}")
