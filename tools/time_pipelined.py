"""Diagnostics of batch.register_pairs_pipelined at config 2: python tools/time_pipelined.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from convexadam_b200 import pipeline  # noqa: E402
from convexadam_b200._native import get_lib  # noqa: E402
from convexadam_b200.batch import iter_register_pipelined  # noqa: E402
from convexadam_b200.convex_adam_MIND import convex_adam_pt  # noqa: E402
from convexadam_b200.synth import make_pair  # noqa: E402

lib = get_lib()
dev = torch.device("cuda:0")
f, m = make_pair((192, 160, 256), seed=0)
fp, mp = f.pin_memory(), m.pin_memory()
kw = dict(mind_r=1, mind_d=2, lambda_weight=1.25, grid_sp=2, disp_hw=6, selected_niter=100, selected_smooth=0, grid_sp_adam=2, ic=True)
P = pipeline.Params(**kw)
fd, md = fp.to(dev), mp.to(dev)

# device-only: N pairs back to back on one stream vs alternating over two streams
def device_only(nstreams, n=6):
    streams = [torch.cuda.Stream(dev) for _ in range(nstreams)]
    def run():
        outs = []
        for i in range(n):
            with torch.cuda.stream(streams[i % nstreams]):
                outs.append(pipeline.register_mind(fd, md, P, lib))
        torch.cuda.synchronize()
        return outs
    run(); run()
    t0 = time.perf_counter(); run(); t1 = time.perf_counter()
    return 1e3 * (t1 - t0) / n

for ns in (1, 2, 3):
    print(f"device-only, {ns} stream(s): {device_only(ns):.2f} ms per pair")
for depth in (1, 2, 3):
    for fld in iter_register_pipelined([(fp, mp)] * 4, depth=depth, device=dev, **kw):
        pass
    t0 = time.perf_counter()
    for fld in iter_register_pipelined([(fp, mp)] * 8, depth=depth, device=dev, **kw):
        pass
    t1 = time.perf_counter()
    print(f"iter_register_pipelined depth={depth}: {1e3 * (t1 - t0) / 8:.2f} ms per pair")
for _ in range(3):
    out = convex_adam_pt(fp, mp, dtype=torch.float32, device=dev, **kw)
t0 = time.perf_counter()
for _ in range(6):
    out = convex_adam_pt(fp, mp, dtype=torch.float32, device=dev, **kw)
t1 = time.perf_counter()
print(f"convex_adam_pt sequential: {1e3 * (t1 - t0) / 6:.2f} ms per pair")
x = out
t0 = time.perf_counter(); y = x.copy(); t1 = time.perf_counter()
print(f"host copy of one (H,W,D,3) float64 field: {1e3 * (t1 - t0):.1f} ms")
