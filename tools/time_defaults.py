"""Per-stage device timings in the package-defaults regime (grid_sp=6, disp_hw=4, 80 it.) at 192x160x256, and config 4."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from convexadam_b200 import pipeline
from convexadam_b200._native import get_lib
from convexadam_b200.convex_adam_MIND import convex_adam_pt
from convexadam_b200.synth import make_pair

lib = get_lib(); dev = torch.device("cuda:0")
f, m = make_pair((192, 160, 256), seed=0)
fd, md = f.to(dev), m.to(dev)
P = pipeline.Params(mind_r=1, mind_d=2, lambda_weight=1.25, grid_sp=6, disp_hw=4, selected_niter=80, grid_sp_adam=2, ic=True)


class T:
    def __init__(self): self.ev = []
    def mark(self, name):
        e = torch.cuda.Event(enable_timing=True); e.record(); self.ev.append((name, e))


for rep in range(3):
    t = T(); pipeline.TIMER = t
    t0 = time.perf_counter(); pipeline.register_mind(fd, md, P, lib); t_launch = time.perf_counter(); torch.cuda.synchronize(); t1 = time.perf_counter()
    pipeline.TIMER = None
    st = {}
    for (n0, e0), (n1, e1) in zip(t.ev[:-1], t.ev[1:]):
        st[n1] = st.get(n1, 0.0) + e0.elapsed_time(e1)
    print(f"rep {rep}: total {1e3 * (t1 - t0):.2f} ms (host launch {1e3 * (t_launch - t0):.2f} ms)", {k: round(v, 2) for k, v in st.items()})
fp, mp = f.pin_memory(), m.pin_memory()
kw = dict(mind_r=1, mind_d=2, lambda_weight=1.25, grid_sp=6, disp_hw=4, selected_niter=80, selected_smooth=0, grid_sp_adam=2, ic=True, dtype=torch.float32, device=dev)
for _ in range(4):
    t0 = time.perf_counter(); convex_adam_pt(fp, mp, **kw); print(f"e2e {1e3 * (time.perf_counter() - t0):.2f} ms")
