"""Adam stage at config 2 (real MIND features of the synthetic pair, start field = the convex stage's) under a list of
tuning-knob settings; every setting runs in its own process because the library reads its knobs once.
    python tools/time_adam_knobs.py                 # the sweep
    python tools/time_adam_knobs.py --one           # one measurement with the current environment (used by the sweep)
Prints ms per 100 iterations and a checksum of the final field (the knobs must not change it, except CA_ADAM_GRAD)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SWEEP = [
    {"CA_ADAM_GRAD": "1"},
    {},
    {"CA_ADAM_BX": "8", "CA_ADAM_BY": "4", "CA_ADAM_BZ": "4"},
    {"CA_ADAM_BX": "16", "CA_ADAM_BY": "4", "CA_ADAM_BZ": "2"},
    {"CA_ADAM_BX": "16", "CA_ADAM_BY": "2", "CA_ADAM_BZ": "4"},
    {"CA_ADAM_BX": "8", "CA_ADAM_BY": "8", "CA_ADAM_BZ": "2"},
    {"CA_ADAM_BX": "8", "CA_ADAM_BY": "2", "CA_ADAM_BZ": "8"},
    {"CA_ADAM_BX": "4", "CA_ADAM_BY": "8", "CA_ADAM_BZ": "4"},
    {"CA_ADAM_BX": "32", "CA_ADAM_BY": "2", "CA_ADAM_BZ": "2"},
    {"CA_ADAM_BX": "32", "CA_ADAM_BY": "1", "CA_ADAM_BZ": "4"},
]


def one():
    import torch
    from convexadam_b200 import pipeline
    from convexadam_b200._native import get_lib
    from convexadam_b200.synth import make_pair

    lib = get_lib()
    dev = torch.device("cuda:0")
    f, m = make_pair((192, 160, 256), seed=0)
    fd, md = f.to(dev), m.to(dev)
    P = pipeline.Params(mind_r=1, mind_d=2, lambda_weight=1.25, grid_sp=2, disp_hw=6, selected_niter=100, grid_sp_adam=2, ic=True)
    fs, _ = lib.mindssc_pooled(fd, 1, 2, 2, 0)
    ms, _ = lib.mindssc_pooled(md, 1, 2, 2, 0)
    disp = pipeline.convex_stage(lib, fs, ms, (192, 160, 256), P)
    out = pipeline.adam_stage(lib, disp, fs, ms, (192, 160, 256), P)
    torch.cuda.synchronize()
    reps = 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = pipeline.adam_stage(lib, disp, fs, ms, (192, 160, 256), P)
    e1.record()
    torch.cuda.synchronize()
    knobs = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith(("CA_ADAM", "CA_BOX3")))
    print(f"adam stage {e0.elapsed_time(e1) / reps:7.3f} ms | sum {out.double().sum().item():.9e} absmax {out.abs().max().item():.6f} | {knobs or 'defaults'}",
          flush=True)


if __name__ == "__main__":
    if "--one" in sys.argv:
        one()
    else:
        for kn in SWEEP:
            env = {k: v for k, v in os.environ.items() if not k.startswith(("CA_ADAM", "CA_BOX3"))}
            env.update(kn)
            subprocess.run([sys.executable, os.path.abspath(__file__), "--one"], env=env, check=False)
