"""Strong-scaling timing of ONE pair split into H slabs over the ranks (convexadam_b200/slab.py; BASELINE config 5 is
the 384^3 / disp_hw=8 case that only fits when split).  Launch with torchrun, one rank per GPU:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/bench_slab.py --shape 192 160 256 --hw 6 [--niter 100] [--steps 3]
Prints one JSON line on rank 0 (device time, max over ranks)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from convexadam_b200 import pipeline, slab  # noqa: E402
from convexadam_b200._native import get_lib  # noqa: E402
from convexadam_b200.synth import make_pair  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shape", type=int, nargs=3, default=[192, 160, 256])
ap.add_argument("--hw", type=int, default=6)
ap.add_argument("--niter", type=int, default=100)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--warmup", type=int, default=1)
args = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
dev = torch.device(f"cuda:{local}")
torch.cuda.set_device(dev)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
lib = get_lib()
f, m = make_pair(tuple(args.shape), seed=0)
fd, md = f.to(dev), m.to(dev)
P = pipeline.Params(mind_r=1, mind_d=2, lambda_weight=1.25, grid_sp=2, disp_hw=args.hw, selected_niter=args.niter,
                    selected_smooth=0, grid_sp_adam=2, ic=True)
for _ in range(args.warmup):
    out = slab.register_mind_slab(fd, md, P, lib)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    out = slab.register_mind_slab(fd, md, P, lib)
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"what": "one pair split into H slabs", "shape": args.shape, "disp_hw": args.hw, "niter": args.niter,
                      "n_gpus": world, "ms_per_pair": float(ms.item()), "pairs_per_s": 1e3 / float(ms.item()),
                      "scaling": "strong", "field_absmax": float(out.abs().max()),
                      "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 2**30}))
if world > 1:
    dist.destroy_process_group()
