"""Pair-parallel batch driver (BASELINE.json config 3; the loop the reference's self-configuring scripts run,
self_configuring/convex_run_withconfig.py:78-156, one process per GPU there too via CUDA_VISIBLE_DEVICES :42-43).

Volume pairs are independent, so the batch shards with NO collective on the data path: rank r of a
torch.distributed world registers pairs r, r+W, r+2W, ...  The only communication is the optional gather of the
finished fields to rank 0 (host tensors, any backend).  `register_fn` is injectable so that the sharding and
gather logic is testable on CPU with the gloo backend.
"""
from typing import Callable, List, Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist


def shard_indices(n_pairs: int, rank: int, world: int) -> List[int]:
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return list(range(rank, n_pairs, world))


def _default_register(fixed, moving, **kw):
    from .convex_adam_MIND import convex_adam_pt
    return convex_adam_pt(fixed, moving, **kw)


def register_pairs(pairs: Sequence, gather: bool = True, register_fn: Optional[Callable] = None, **kw):
    """pairs: sequence of (fixed, moving) volumes (or callables returning such a tuple, evaluated only on the owning
    rank).  Returns, on rank 0 (or on every rank when gather=False: only its own entries), a list with one
    (H,W,D,3) float64 field per pair; None for pairs this rank does not hold.  A failing pair raises
    RuntimeError naming the pair index."""
    register_fn = register_fn or _default_register
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    mine = shard_indices(len(pairs), rank, world)
    results: List[Optional[np.ndarray]] = [None] * len(pairs)
    for i in mine:
        item = pairs[i]() if callable(pairs[i]) else pairs[i]
        try:
            results[i] = register_fn(item[0], item[1], **kw)
        except Exception as e:  # surface the pair index, no recovery semantics
            raise RuntimeError(f"registration of pair {i} failed on rank {rank}: {e}") from e
    if not gather or world == 1:
        return results
    gathered = [None] * world if rank == 0 else None
    dist.gather_object({i: results[i] for i in mine}, gathered, dst=0)
    if rank != 0:
        return results
    for part in gathered:
        for i, f in part.items():
            results[i] = f
    return results
