"""debug aid: per-phase cycles of the producer / consumer correlate kernel (needs convexadam_b200/variants/lib_cpprof.so,
tools/build_cp_profile.sh)"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from convexadam_b200._native import Lib
from convexadam_b200.synth import make_pair
lib = Lib(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "convexadam_b200", "variants", "lib_cpprof.so"))
dev = torch.device("cuda:0")
f, m = make_pair((192, 160, 256), seed=0)
fs, _ = lib.mindssc_pooled(f.to(dev), 1, 2, 2); ms_, _ = lib.mindssc_pooled(m.to(dev), 1, 2, 2)
ssd = torch.empty((2197, 96, 80, 128), dtype=torch.float32, device=dev)
lib.correlate(fs, ms_, 6, out=ssd, fast=True); torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * 24)()
lib.dll.ca_debug_cp_prof(buf, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); lib.correlate(fs, ms_, 6, out=ssd, fast=True); e1.record(); torch.cuda.synchronize()
lib.dll.ca_debug_cp_prof(buf, 0)
print(f"kernel {e0.elapsed_time(e1):.3f} ms; hand-off error flag {lib.dll.ca_debug_corr_pipe_error(1)}")
nprod, nstep, ncons = max(1, buf[4]), max(1, buf[12]), max(1, buf[14])
print(f"producer CTAs {nprod}: lifetime {buf[5] / nprod:.0f} cycles = wait tiles {buf[0] / nprod:.0f} + arithmetic (incl. slot wait) {buf[1] / nprod:.0f}"
      f" [slot wait {buf[2] / nprod:.0f}] + fence {buf[3] / nprod:.0f}; loader: tiles issue + ring-slot poll {buf[6] / nprod:.0f}")
print(f"consumer CTAs {ncons}: lifetime {buf[13] / ncons:.0f} cycles; per plane step ({nstep} steps): wait full {buf[8] / nstep:.0f}, phase A {buf[9] / nstep:.0f},"
      f" plane barrier {buf[10] / nstep:.0f}, phase B + T_H + store {buf[11] / nstep:.0f}")
print(f"consumer loader per plane: wait empty {buf[16] / nstep:.0f}, poll producers {buf[17] / nstep:.0f}, fence + issue {buf[18] / nstep:.0f},"
      f" wait landed + signal {buf[19] / nstep:.0f}")
