import os, sys, torch
sys.path.insert(0, os.getcwd())
from convexadam_b200 import pipeline
from convexadam_b200._native import get_lib
from convexadam_b200.synth import make_pair
lib = get_lib(); dev = torch.device("cuda:0")
f, m = make_pair((192, 160, 256), seed=0); f = f.to(dev); m = m.to(dev)
ssd = torch.empty((2197, 96, 80, 128), dtype=torch.float32, device=dev)
for ov in ("0", "1"):
    P = pipeline.Params(selected_niter=100, overlap=ov, grid_sp=2, disp_hw=6, grid_sp_adam=2)
    for _ in range(3): pipeline.register_mind(f, m, P, lib, ssd)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): pipeline.register_mind(f, m, P, lib, ssd)
    e1.record(); torch.cuda.synchronize()
    print("pipe", os.environ.get("CA_CORR_PIPE", "0"), os.environ.get("CA_PIPE_NW",""), os.environ.get("CA_PIPE_PB",""), os.environ.get("CA_PIPE_R",""), "overlap", ov, "ms/pair", e0.elapsed_time(e1) / 5, "err", lib.dll.ca_debug_corr_pipe_error(1))
