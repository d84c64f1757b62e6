"""N>1 host logic of the slab-parallel path (convexadam_b200/slab.py) on CPU: world size 2, gloo, the kernels being
the CPU-emulation build of the same sources.  The split result must equal the single-process pipeline bit for bit.
(The same comparison runs on 2 real GPUs over NCCL in tests/test_slab_gpu.py.)"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from convexadam_b200 import pipeline, slab
from convexadam_b200._native import Lib
from tests.emu.build_emu import build


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, shape, kw, path, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from convexadam_b200.synth import make_pair
        lib = Lib(path, expect_cuda=False)
        f, m = make_pair(shape, seed=11, disp_scale=0.15)
        got = slab.register_mind_slab(f, m, pipeline.Params(**kw), lib)
        np.save(os.path.join(out_dir, f"rank{rank}.npy"), got.numpy())
    finally:
        dist.destroy_process_group()


def test_split_planes():
    assert slab.split_planes(10, 3) == [(0, 4), (4, 7), (7, 10)]
    assert slab.split_planes(8, 8) == [(i, i + 1) for i in range(8)]
    with pytest.raises(ValueError):
        slab.split_planes(3, 4)


@pytest.mark.parametrize("shape,niter,mode", [((32, 12, 16), 3, ""),            # 8 Adam planes per rank: 7-plane halo exchange
                                               ((20, 12, 16), 2, ""),            # 5 per rank: all-gather fallback
                                               ((32, 12, 16), 2, "certified")])  # the certified pieces on slabs
def test_two_ranks_equal_single_process(tmp_path, shape, niter, mode):
    path = build()
    kw = dict(mind_r=1, mind_d=2, lambda_weight=1.25, grid_sp=2, disp_hw=2, selected_niter=niter, selected_smooth=0,
              grid_sp_adam=2, ic=True, mode=mode)
    from convexadam_b200.synth import make_pair
    lib = Lib(path, expect_cuda=False)
    f, m = make_pair(shape, seed=11, disp_scale=0.15)
    want = pipeline.register_mind(f, m, pipeline.Params(**kw), lib).numpy()
    single = slab.register_mind_slab(f, m, pipeline.Params(**kw), lib).numpy()      # world size 1 through the slab entry points
    assert (single.view(np.uint32) == want.view(np.uint32)).all()
    mp.spawn(_worker, args=(2, _free_port(), shape, kw, path, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        got = np.load(tmp_path / f"rank{r}.npy")
        assert got.shape == want.shape
        assert (got.view(np.uint32) == want.view(np.uint32)).all(), f"rank {r}: max|diff| {np.abs(got - want).max()}"
