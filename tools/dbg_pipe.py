"""debug aid: one certified-mode correlate call (producer / consumer kernel, csrc/correlate_pipe.cu) on a small shape,
synchronised, compared with the exact volume; prints the kernel's hand-off error flag:
    python tools/dbg_pipe.py C h w d hw"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from convexadam_b200._native import get_lib
lib = get_lib()
dev = torch.device("cuda:0")
shape = tuple(int(v) for v in sys.argv[1:5]) if len(sys.argv) > 4 else (12, 7, 21, 40)
hw = int(sys.argv[5]) if len(sys.argv) > 5 else 2
r = np.random.RandomState(0)
fix = torch.from_numpy(r.rand(*shape).astype(np.float32)).to(dev); mov = torch.from_numpy(r.rand(*shape).astype(np.float32)).to(dev)
n0 = lib.launch_count()
a, _ = lib.correlate(fix, mov, hw, fast=True)
torch.cuda.synchronize()
print("fast ok; launches", lib.launch_count() - n0, "max", float(a.max()), "hand-off error flag", lib.dll.ca_debug_corr_pipe_error(1))
b, _ = lib.correlate(fix, mov, hw, fast=False, want_argmin=False)
torch.cuda.synchronize()
rel = ((a - b).abs() / b.clamp_min(1e-30)).max()
print("fast vs exact rel", float(rel), "bound", 98 * 2 ** -24)
