"""CUDA-event timings of the convex stage's two calls at the BASELINE config-2 shape, certified and exact mode (tuning
aid; the env knobs CA_RAW_MINB / CA_BOX_THREADS / CA_CORR_CHUNK are read by the library once per process):
    python tools/time_corr.py [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from convexadam_b200._native import get_lib  # noqa: E402
from convexadam_b200.synth import make_pair  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
if "--lib" in sys.argv:          # a variant build of the library (tuning): python tools/time_corr.py 5 --lib path/to/lib.so
    from convexadam_b200._native import Lib
    lib = Lib(sys.argv[sys.argv.index("--lib") + 1])
    print("library:", lib.path)
else:
    lib = get_lib()
dev = torch.device("cuda:0")
f, m = make_pair((192, 160, 256), seed=0)
fs, _ = lib.mindssc_pooled(f.to(dev), 1, 2, 2)
ms_, _ = lib.mindssc_pooled(m.to(dev), 1, 2, 2)
ssd = torch.empty((2197, 96, 80, 128), dtype=torch.float32, device=dev)
mesh = lib.disp_mesh(6, dev)


def timeit(fn, n=reps):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        r = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, r


knobs = " ".join(f"{k}={os.environ[k]}" for k in ("CA_RAW_MINB", "CA_BOX_THREADS", "CA_CORR_CHUNK", "CA_CORR_PIPE", "CA_PIPE_NW",
                                                  "CA_PIPE_R", "CA_PIPE_PB", "CA_CORR_FUSED") if k in os.environ)
t_fast, _ = timeit(lambda: lib.correlate(fs, ms_, 6, out=ssd, fast=True))
t_cert, _ = timeit(lambda: lib.coupled_convex_certified(ssd, fs, ms_, mesh, 6))
line = f"[{knobs}] certified: correlate_fast/dir {t_fast:.3f} ms, coupled_certified/dir {t_cert:.3f} ms"
if "--exact" in sys.argv:
    t_ex, _ = timeit(lambda: lib.correlate(fs, ms_, 6, out=ssd, want_argmin=False))
    t_cc, _ = timeit(lambda: lib.coupled_convex(ssd, None, mesh, 6))
    line += f" | exact: correlate/dir {t_ex:.3f} ms, coupled/dir {t_cc:.3f} ms"
try:
    line += f" | pipe hand-off error flag {lib.dll.ca_debug_corr_pipe_error(1)}"
except AttributeError:
    pass
print(line)
