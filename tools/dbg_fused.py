"""debug aid: one fused correlate call on a small shape, synchronised, compared with the exact volume"""
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
os.environ.setdefault("CA_CORR_FUSED", "1")
a, _ = lib.correlate(fix, mov, hw, fast=True)
torch.cuda.synchronize()
print("fused ok; max", float(a.max()))
b, _ = lib.correlate(fix, mov, hw, fast=False, want_argmin=False)
torch.cuda.synchronize()
rel = ((a - b).abs() / b.clamp_min(1e-30)).max()
print("fused vs exact rel", float(rel), "bound", 98 * 2 ** -24)
