"""Slab-parallel path on real GPUs over NCCL: one pair split over 2 ranks must equal the single-GPU result bit for
bit.  Needs >= 2 GPUs (skipped otherwise; run with `gpurun --gpus 2 -- python -m pytest tests/test_slab_gpu.py -m gpu`)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, shape, kw, out_dir):
    import torch.distributed as dist
    from convexadam_b200.convex_adam_MIND import convex_adam_pt
    from convexadam_b200.synth import make_pair
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    try:
        f, m = make_pair(shape, seed=5)
        got = convex_adam_pt(f, m, dtype=torch.float32, device=torch.device(f"cuda:{rank}"), split_over=True, **kw)
        np.save(os.path.join(out_dir, f"rank{rank}.npy"), got)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpus_equal_one(tmp_path):
    import torch.multiprocessing as mp
    from convexadam_b200.convex_adam_MIND import convex_adam_pt
    from convexadam_b200.synth import make_pair
    shape = (96, 80, 64)
    kw = dict(mind_r=1, mind_d=2, lambda_weight=1.25, grid_sp=2, disp_hw=4, selected_niter=12, selected_smooth=0,
              grid_sp_adam=2, ic=True)
    f, m = make_pair(shape, seed=5)
    want = convex_adam_pt(f, m, dtype=torch.float32, device=torch.device("cuda:0"), **kw)
    mp.spawn(_worker, args=(2, _free_port(), shape, kw, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        got = np.load(tmp_path / f"rank{r}.npy")
        assert (got == want).all(), f"rank {r}: max|diff| {np.abs(got - want).max()}"
