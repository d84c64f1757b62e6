"""CUDA-event timing of 100 Adam iterations at the config-2 grid for a given build of the library:
    python tools/time_adam.py [path/to/lib.so]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from convexadam_b200._native import LIB_PATH, Lib  # noqa: E402

path = sys.argv[1] if len(sys.argv) > 1 else LIB_PATH
lib = Lib(path)
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
fs = torch.rand(12, 96, 80, 128, generator=g).to(dev)
ms = torch.rand(12, 96, 80, 128, generator=g).to(dev)
p = (torch.randn(3, 96, 80, 128, generator=g) * 2).to(dev)
m, v = torch.zeros_like(p), torch.zeros_like(p)
lib.adam_run(fs, ms, 1.25, p, m, v, 1, 10)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    lib.adam_run(fs, ms, 1.25, p, m, v, 1, 100)
e1.record(); torch.cuda.synchronize()
print(f"{os.path.basename(path)}: adam100 {e0.elapsed_time(e1) / 3:.3f} ms")
