"""One full-size registration (BASELINE config 2 shape) for ncu captures: python tools/profile_pair.py [niter] [shape...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from convexadam_b200 import pipeline  # noqa: E402
from convexadam_b200._native import get_lib  # noqa: E402
from convexadam_b200.synth import make_pair  # noqa: E402

niter = int(sys.argv[1]) if len(sys.argv) > 1 else 2
shape = tuple(int(v) for v in sys.argv[2:5]) if len(sys.argv) >= 5 else (192, 160, 256)
lib = get_lib()
dev = torch.device("cuda:0")
f, m = make_pair(shape, seed=0)
P = pipeline.Params(mind_r=1, mind_d=2, lambda_weight=1.25, grid_sp=2, disp_hw=6, selected_niter=niter, selected_smooth=0,
                    grid_sp_adam=2, ic=True)
out = pipeline.register_mind(f.to(dev), m.to(dev), P, lib)
torch.cuda.synchronize()
print("done", tuple(out.shape), float(out.abs().max()), "launches", lib.launch_count())
