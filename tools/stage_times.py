"""Per-stage CUDA-event timings at the BASELINE config-2 shape (tuning aid; env knobs CA_RAW_MINB / CA_CORR_CHUNK are
read by the library once per process):  python tools/stage_times.py [reps]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from convexadam_b200 import pipeline  # noqa: E402
from convexadam_b200._native import get_lib  # noqa: E402
from convexadam_b200.convex_adam_MIND import convex_adam_pt  # noqa: E402
from convexadam_b200.synth import make_pair  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
lib = get_lib()
dev = torch.device("cuda:0")
f, m = make_pair((192, 160, 256), seed=0)
fd, md = f.to(dev), m.to(dev)


def timeit(fn, n=reps):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        r = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, r


t_mind, (fs, _) = timeit(lambda: lib.mindssc_pooled(fd, 1, 2, 2))
ms_, _ = lib.mindssc_pooled(md, 1, 2, 2)
ssd = torch.empty((2197, 96, 80, 128), dtype=torch.float32, device=dev)
t_corr, (_, am) = timeit(lambda: lib.correlate(fs, ms_, 6, out=ssd))
t_corr_noarg, _ = timeit(lambda: lib.correlate(fs, ms_, 6, out=ssd, want_argmin=False))
mesh = lib.disp_mesh(6, dev)
t_cc, soft = timeit(lambda: lib.coupled_convex(ssd, am, mesh, 6))
p = torch.zeros(3, 96, 80, 128, device=dev)
mm, vv = torch.zeros_like(p), torch.zeros_like(p)
t_adam, _ = timeit(lambda: lib.adam_run(fs, ms_, 1.25, p, mm, vv, 1, 100), n=max(1, reps // 2))
print(f"CA_RAW_MINB={os.environ.get('CA_RAW_MINB')} CA_CORR_CHUNK={os.environ.get('CA_CORR_CHUNK')}: mind/img {t_mind:.3f} ms | "
      f"correlate/dir {t_corr:.3f} ms (without argmin {t_corr_noarg:.3f}) | coupled/dir {t_cc:.3f} ms | adam100 {t_adam:.3f} ms")

if "--e2e" in sys.argv:
    fp, mp = f.pin_memory(), m.pin_memory()
    kw = dict(mind_r=1, mind_d=2, lambda_weight=1.25, grid_sp=2, disp_hw=6, selected_niter=100, selected_smooth=0, grid_sp_adam=2,
              ic=True, dtype=torch.float32, device=dev)
    for _ in range(3):
        convex_adam_pt(fp, mp, **kw)
    t0 = time.perf_counter(); out = convex_adam_pt(fp, mp, **kw); t1 = time.perf_counter()
    print(f"e2e convex_adam_pt: {1e3 * (t1 - t0):.2f} ms")
    t0 = time.perf_counter(); a = fp.to(dev, non_blocking=True); b = mp.to(dev, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"  h2d 2 images: {1e3 * (t1 - t0):.2f} ms")
    P = pipeline.Params(mind_r=1, mind_d=2, lambda_weight=1.25, grid_sp=2, disp_hw=6, selected_niter=100, grid_sp_adam=2, ic=True)
    t0 = time.perf_counter(); disp = pipeline.register_mind(a, b, P, lib); t_launch = time.perf_counter(); torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"  pipeline: {1e3 * (t1 - t0):.2f} ms (host launch part {1e3 * (t_launch - t0):.2f} ms)")
    t0 = time.perf_counter(); packed = lib.pack_field(disp); torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"  pack f64: {1e3 * (t1 - t0):.2f} ms")
    t0 = time.perf_counter(); host = torch.empty(packed.shape, dtype=torch.float64, pin_memory=True); t1 = time.perf_counter()
    print(f"  pinned alloc: {1e3 * (t1 - t0):.2f} ms")
    t0 = time.perf_counter(); host.copy_(packed, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"  d2h 189 MB: {1e3 * (t1 - t0):.2f} ms")
    del host
    t0 = time.perf_counter(); host = torch.empty(packed.shape, dtype=torch.float64, pin_memory=True); t1 = time.perf_counter()
    print(f"  pinned alloc (2nd): {1e3 * (t1 - t0):.2f} ms")
