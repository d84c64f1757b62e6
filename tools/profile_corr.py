"""One correlate + coupled_convex call at the BASELINE config-2 shape (for ncu captures): python tools/profile_corr.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from convexadam_b200._native import get_lib  # noqa: E402
from convexadam_b200.synth import make_pair  # noqa: E402

lib = get_lib()
dev = torch.device("cuda:0")
f, m = make_pair((192, 160, 256), seed=0)
fs, _ = lib.mindssc_pooled(f.to(dev), 1, 2, 2)
ms, _ = lib.mindssc_pooled(m.to(dev), 1, 2, 2)
if "--fast" in sys.argv:       # the certified path: approximate volume + certified selections
    ssd, _ = lib.correlate(fs, ms, 6, fast=True)
    if "--cc" in sys.argv:
        soft = lib.coupled_convex_certified(ssd, fs, ms, lib.disp_mesh(6, dev), 6)
else:
    ssd, am = lib.correlate(fs, ms, 6)
    if "--cc" in sys.argv:
        soft = lib.coupled_convex(ssd, am, lib.disp_mesh(6, dev), 6)
torch.cuda.synchronize()
print("done", lib.launch_count())
