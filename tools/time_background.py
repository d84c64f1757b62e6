"""Certified mode on a pair with a large constant background (CT-like: air outside the body): flagged counts and times.
python tools/time_background.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from convexadam_b200._native import get_lib  # noqa: E402
from convexadam_b200.synth import make_pair  # noqa: E402

lib = get_lib()
dev = torch.device("cuda:0")
shape = (192, 160, 256)
f, m = make_pair(shape, seed=0)
xs = [torch.linspace(-1, 1, s) for s in shape]
r2 = xs[0][:, None, None] ** 2 + xs[1][None, :, None] ** 2 + xs[2][None, None, :] ** 2
f[r2 > 0.6] = -1000.0
m[r2 > 0.55] = -1000.0
fs, _ = lib.mindssc_pooled(f.to(dev), 1, 2, 2)
ms, _ = lib.mindssc_pooled(m.to(dev), 1, 2, 2)
mesh = lib.disp_mesh(6, dev)
ssd = torch.empty((2197, 96, 80, 128), dtype=torch.float32, device=dev)


def timeit(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        r = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, r


t_e, _ = timeit(lambda: lib.correlate(fs, ms, 6, out=ssd, want_argmin=False))
t_ce, (soft_e, sel_e) = timeit(lambda: lib.coupled_convex(ssd, None, mesh, 6, want_sel=True))
t_f, _ = timeit(lambda: lib.correlate(fs, ms, 6, out=ssd, fast=True))
t_cf, (soft_c, sel_c, fl) = timeit(lambda: lib.coupled_convex_certified(ssd, fs, ms, mesh, 6, want_sel=True, want_flagged=True))
print(f"constant background ({float((r2 > 0.6).float().mean()) * 100:.0f} % of the fixed image): exact {t_e:.2f} + {t_ce:.2f} ms | certified {t_f:.2f} + {t_cf:.2f} ms")
print("flagged per selection:", fl.tolist(), "| selections equal:", bool((sel_c[1:] == sel_e).all()),
      "| soft bits equal:", bool((soft_c.view(torch.int32) == soft_e.view(torch.int32)).all()))
