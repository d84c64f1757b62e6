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


_STREAMS = {}
_RINGS = {}


def _pipeline_streams(dev, depth):
    key = str(dev)
    have = _STREAMS.setdefault(key, [])
    while len(have) < depth:
        have.append(torch.cuda.Stream(dev))        # cached: torch's allocator keeps one memory pool per stream
    return have[:depth]


def iter_register_pipelined(pairs: Sequence, depth: int = 2, device=None, dtype=torch.float32, copy: bool = False, **kw):
    """Generator over the fields of `pairs`, in order, with `depth` pairs in flight on as many CUDA streams: the Adam
    stage of one pair (small latency-bound kernels) shares the GPU with the cost-volume stage of the next, and the
    host<->device copies of one pair overlap the compute of another.  The kernels are the ones `convex_adam_pt` runs,
    each pair on its own stream, so every field is bit-identical to `convex_adam_pt(fixed, moving, dtype=dtype, **kw)`.
    pairs: sequence of (fixed, moving) volumes (or callables returning such a tuple).  Yields (H,W,D,3) float64 arrays.
    copy=False (default): each array is a view of one of depth+2 pinned staging buffers that are reused round-robin --
    it stays valid until TWO further fields have been yielded (consume or copy it before asking for the third); copying
    189 MB per field on the host would cost more than the registration itself.  copy=True yields arrays that own their memory.
    MIND path, no masks."""
    from . import pipeline
    from ._native import get_lib
    from .convex_adam_utils import validate_image
    if depth < 1:
        raise ValueError("depth must be >= 1")
    lib = get_lib()
    dev = torch.device(device if device is not None else "cuda")
    names = ("mind_r", "mind_d", "lambda_weight", "grid_sp", "disp_hw", "selected_niter", "selected_smooth", "grid_sp_adam", "ic")
    unknown = set(kw) - set(names)
    if unknown:
        raise TypeError(f"unexpected arguments {sorted(unknown)}")
    # the pairs in flight already share the GPU: each runs its two directions one after the other (depth x 1 cost volumes)
    prm = pipeline.Params(overlap="0" if depth > 1 else "", **{k: kw[k] for k in names if k in kw})
    with torch.cuda.device(dev):
        main = torch.cuda.current_stream(dev)
        streams = _pipeline_streams(dev, depth)
        pending = []                                   # (host tensor, event) in issue order
        ring = _RINGS.setdefault(str(dev), {})         # pinned staging buffers, allocated once per device and shape
        nring = depth + 2
        for i in range(len(pairs)):
            item = pairs[i]() if callable(pairs[i]) else pairs[i]
            fixed, moving = validate_image(item[0]).float(), validate_image(item[1]).float()
            st = streams[i % depth]
            st.wait_stream(main)
            try:
                with torch.cuda.stream(st):
                    fd = fixed.to(dev, non_blocking=True).contiguous()
                    md = moving.to(dev, non_blocking=True).contiguous()
                    disp = pipeline.register_mind(fd, md, prm, lib)
                    packed = lib.pack_field(disp, round_half=(dtype == torch.float16))
                    key = (tuple(packed.shape), i % nring)
                    host = ring.get(key)
                    if host is None:
                        host = ring[key] = torch.empty(packed.shape, dtype=torch.float64, pin_memory=True)
                    host.copy_(packed, non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(st)
            except Exception as e:
                raise RuntimeError(f"registration of pair {i} failed: {e}") from e
            pending.append((host, ev))
            del fd, md, disp, packed, host
            while len(pending) >= depth:               # at most `depth` pairs enqueued ahead of the consumer
                h, e = pending.pop(0)
                e.synchronize()
                yield h.numpy().copy() if copy else h.numpy()
        while pending:
            h, e = pending.pop(0)
            e.synchronize()
            yield h.numpy().copy() if copy else h.numpy()
        for st in streams:
            main.wait_stream(st)


def register_pairs_pipelined(pairs: Sequence, depth: int = 2, device=None, dtype=torch.float32, **kw):
    """list(iter_register_pipelined(..., copy=True)): every field is kept, each owning its memory."""
    return list(iter_register_pipelined(pairs, depth=depth, device=device, dtype=dtype, copy=True, **kw))
