"""MIND-SSC + pooling per image at config 2, both forms of phase A (CA_MIND_ROWS is read per call), GPU warmed up first:
    python tools/time_mind.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from convexadam_b200._native import get_lib  # noqa: E402
from convexadam_b200.synth import make_pair  # noqa: E402

lib = get_lib()
dev = torch.device("cuda:0")
f, _ = make_pair((192, 160, 256), seed=0)
fd = f.to(dev)
for _ in range(200):
    lib.mindssc_pooled(fd, 1, 2, 2)
torch.cuda.synchronize()
for rnd in range(2):
    for rows in ("1", "4"):
        os.environ["CA_MIND_ROWS"] = rows
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            lib.mindssc_pooled(fd, 1, 2, 2)
        e1.record()
        torch.cuda.synchronize()
        print(f"CA_MIND_ROWS={rows}: {e0.elapsed_time(e1) / 50:.3f} ms per image (phase A + stats + phase B)")
