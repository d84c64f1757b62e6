"""debug aid: per-phase cycles of the fused correlate kernel (needs convexadam_b200/variants/lib_cfprof.so)"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["CA_CORR_FUSED"] = "1"
import torch
from convexadam_b200._native import Lib
from convexadam_b200.synth import make_pair
lib = Lib(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "convexadam_b200", "variants", "lib_cfprof.so"))
dev = torch.device("cuda:0")
f, m = make_pair((192, 160, 256), seed=0)
fs, _ = lib.mindssc_pooled(f.to(dev), 1, 2, 2); ms_, _ = lib.mindssc_pooled(m.to(dev), 1, 2, 2)
ssd = torch.empty((2197, 96, 80, 128), dtype=torch.float32, device=dev)
lib.correlate(fs, ms_, 6, out=ssd, fast=True); torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * 8)()
lib.dll.ca_debug_cf_prof(buf, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); lib.correlate(fs, ms_, 6, out=ssd, fast=True); e1.record(); torch.cuda.synchronize()
lib.dll.ca_debug_cf_prof(buf, 0)
names = ["raw: wait TMA stage", "raw: channel arithmetic", "raw: stage release + refill issue", "raw: wait exchange free", "raw: exchange write",
         "owner: wait exchange full", "owner: filter + store"]
tot_raw = sum(buf[i] for i in range(5)); tot_own = buf[5] + buf[6]
print(f"kernel {e0.elapsed_time(e1):.3f} ms")
for i, n in enumerate(names):
    t = max(1, tot_raw if i < 5 else tot_own)
    print(f"  {n:36s} {buf[i] / 1e6:10.1f} Mcycles  {100.0 * buf[i] / t:5.1f} % of the role")
