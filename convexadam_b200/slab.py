"""One oversized pair split into H slabs over the ranks of a torch.distributed group (BASELINE.json config 5,
SURVEY.md §8(e)): each rank holds 1/W of the cost volume -- the only array of the path that does not fit one GPU
(139 GB per direction at 384^3, disp_hw=8) -- and 1/W of the Adam state.

What is exchanged (NCCL over NVLink on the GPU box; gloo in the CPU tests), nothing else crosses ranks:

  * coupled_convex (reference utils.py:93-109): before each of the 7 smoothing steps `box3(mesh[sel])` one plane of
    selected labels (w*d int64) goes to each neighbouring slab;
  * after the convex stage the two coarse fields are all-gathered (3*h*w*d floats each); inverse consistency and
    upsampling are replicated (they are <1 % of the work);
  * Adam (MIND.py:147-181): a rank owns the planes [a0, a1) of the Adam grid and computes on [a0-7, a1+7): one
    iteration is exact 7 planes inside a cut (3 for u = B^3 p, 1 for the regulariser's neighbours, 3 for B^3 g_u),
    so before every iteration the 7-plane margins of p are refreshed from the neighbours' owned planes;
  * the final u (owned planes) is all-gathered and upsampled on every rank.

The features are computed redundantly on every rank (2 x 12 x h*w*d floats, MIND is ~3 % of the work) so that the
correlation kernel needs no exchange at all.  Every piece is the same kernel as the single-GPU path restricted to a
window, and windows reproduce the full computation bit for bit (tests/kernel_checks.py: check_correlate_slab,
check_cc_pieces, check_adam_window), hence the split result equals the single-GPU result exactly.
"""
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

from . import pipeline
from ._native import Lib, get_lib

ADAM_MARGIN = 7          # planes of context one Adam iteration needs across a cut (3 + 1 + 3)
CC_COEFFS = (0.003, 0.01, 0.03, 0.1, 0.3, 1.0)   # utils.py:98


def split_planes(n: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous, balanced [begin, end) ranges of n planes for `world` ranks (the first n % world get one more)."""
    if world < 1 or n < world:
        raise ValueError(f"cannot split {n} planes over {world} ranks")
    base, extra = divmod(n, world)
    out, b = [], 0
    for r in range(world):
        e = b + base + (1 if r < extra else 0)
        out.append((b, e))
        b = e
    return out


class _Comm:
    """Neighbour exchange and all-gather along the split axis, over a torch.distributed group (or nothing when W=1)."""

    def __init__(self, group=None):
        self.group = group
        self.on = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if self.on else 1
        self.rank = dist.get_rank(group) if self.on else 0

    def _peer(self, r):
        return dist.get_global_rank(self.group, r) if self.group is not None else r

    def exchange(self, to_lo: Optional[torch.Tensor], to_hi: Optional[torch.Tensor]):
        """Sends `to_lo` to rank-1 and `to_hi` to rank+1; returns (from rank-1, from rank+1), None at the ends.
        Both neighbours send tensors of the same shape as the ones this rank sends in the opposite direction."""
        if self.world == 1:
            return None, None
        ops, from_lo, from_hi = [], None, None
        if self.rank > 0:
            from_lo = torch.empty(to_lo.shape, dtype=to_lo.dtype, device=to_lo.device)
            ops.append(dist.P2POp(dist.isend, to_lo.contiguous(), self._peer(self.rank - 1), self.group))
            ops.append(dist.P2POp(dist.irecv, from_lo, self._peer(self.rank - 1), self.group))
        if self.rank < self.world - 1:
            from_hi = torch.empty(to_hi.shape, dtype=to_hi.dtype, device=to_hi.device)
            ops.append(dist.P2POp(dist.isend, to_hi.contiguous(), self._peer(self.rank + 1), self.group))
            ops.append(dist.P2POp(dist.irecv, from_hi, self._peer(self.rank + 1), self.group))
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        return from_lo, from_hi

    def gather_planes(self, x: torch.Tensor, ranges: List[Tuple[int, int]], dim: int = 1) -> torch.Tensor:
        """All-gather of per-rank plane ranges along `dim` (ranges may differ by one plane: padded, then trimmed)."""
        if self.world == 1:
            return x
        mx = max(e - b for b, e in ranges)
        pad = mx - x.shape[dim]
        xs = x.contiguous()
        if pad:
            shape = list(x.shape)
            shape[dim] = pad
            xs = torch.cat([xs, xs.new_zeros(shape)], dim).contiguous()
        parts = [torch.empty_like(xs) for _ in range(self.world)]
        dist.all_gather(parts, xs, group=self.group)
        return torch.cat([p.narrow(dim, 0, e - b) for p, (b, e) in zip(parts, ranges)], dim).contiguous()


def coupled_convex_slab(lib: Lib, comm: _Comm, ssd, mesh, disp_hw, certified=None):
    """argmin (utils.py:87) + coupled_convex (utils.py:93-109) on this rank's slab: ssd (L,hs,w,d) -> disp_soft
    (3,hs,w,d).  One full pass builds the block minima and the argmin; the six regularised passes are pruned.
    certified = (fix, mov, x_begin): ssd is the approximate volume of correlate_slab(fast=True) and every selection is
    certified against the features (same selections as the exact volume, bit for bit)."""
    _, hs, w, d = ssd.shape
    if certified is not None:
        cf, cm, cx0 = certified
        bmin, sel = lib.cc_blockmin_certified(ssd, disp_hw, mesh, cf, cm, cx0)
    else:
        bmin, sel = lib.cc_blockmin(ssd, disp_hw)
    outside = torch.full((1, w, d), -1, dtype=torch.int64, device=ssd.device)
    soft = None
    for j in range(len(CC_COEFFS) + 1):
        lo, hi = comm.exchange(sel[:1], sel[-1:])
        window = torch.cat([lo if lo is not None else outside, sel, hi if hi is not None else outside], 0).contiguous()
        soft = lib.cc_smooth(window, mesh, disp_hw)                                   # utils.py:96 / :107
        if j == len(CC_COEFFS):
            break
        if certified is not None:
            sel = lib.cc_select_certified(ssd, disp_hw, mesh, soft, CC_COEFFS[j], bmin, sel, cf, cm, cx0)
        else:
            sel = lib.cc_select(ssd, disp_hw, mesh, soft, CC_COEFFS[j], bmin=bmin, prev=sel)   # utils.py:104-105
    return soft


def adam_slab(lib: Lib, comm: _Comm, p_full, fix_a, mov_a, prm: pipeline.Params):
    """The Adam loop with the grid split into H slabs.  p_full: (3,ha,wa,da) start parameters (replicated).
    Returns u of the last forward, (3,ha,wa,da), on every rank."""
    _, ha, wa, da = p_full.shape
    ranges = split_planes(ha, comm.world)
    a0, a1 = ranges[comm.rank]
    M = ADAM_MARGIN
    halo_ok = all(e - b >= M for b, e in ranges)      # otherwise a margin would span more than one neighbour
    lo, hi = max(0, a0 - M), min(ha, a1 + M)
    p = p_full[:, lo:hi].contiguous()
    m = torch.zeros_like(p)
    v = torch.zeros_like(p)
    ws, u = None, None
    for it in range(1, int(prm.selected_niter) + 1):
        if it > 1 and comm.world > 1:
            if halo_ok:
                # my owned planes next to each cut -> the neighbour's margin; its owned planes -> my margin
                from_lo, from_hi = comm.exchange(p[:, a0 - lo:a0 - lo + M], p[:, a1 - lo - M:a1 - lo])
                if from_lo is not None:
                    p[:, :a0 - lo] = from_lo
                if from_hi is not None:
                    p[:, a1 - lo:] = from_hi
            else:
                full = comm.gather_planes(p[:, a0 - lo:a1 - lo], ranges)
                p.copy_(full[:, lo:hi])
        u, ws = lib.adam_window(fix_a, mov_a, float(prm.lambda_weight), lo, hi, p, m, v, it, 1, ws=ws, features_ready=it > 1)
    return comm.gather_planes(u[:, a0 - lo:a1 - lo], ranges)


def register_mind_slab(fixed_dev, moving_dev, prm: pipeline.Params, lib: Lib = None, group=None):
    """Slab-parallel form of pipeline.register_mind: every rank passes the SAME two (H,W,D) float32 images (on its own
    device) and gets the same dense field (3,H,W,D) back.  With a world size of 1 this is the single-GPU pipeline
    run through the slab entry points."""
    lib = lib or get_lib()
    comm = _Comm(group)
    H, W, D = fixed_dev.shape
    same = prm.grid_sp == prm.grid_sp_adam
    need_adam = prm.lambda_weight > 0
    gs_b = 0 if (same or not need_adam) else prm.grid_sp_adam
    fix_s, fix_a = lib.mindssc_pooled(fixed_dev, prm.mind_r, prm.mind_d, prm.grid_sp, gs_b)
    mov_s, mov_a = lib.mindssc_pooled(moving_dev, prm.mind_r, prm.mind_d, prm.grid_sp, gs_b)
    if same:
        fix_a, mov_a = fix_s, mov_s
    return register_features_slab(lib, comm, fix_s, mov_s, fix_a, mov_a, (H, W, D), prm)


def register_features_slab(lib: Lib, comm: _Comm, fix_s, mov_s, fix_a, mov_a, shape, prm: pipeline.Params):
    C, h, w, d = fix_s.shape
    ranges = split_planes(h, comm.world)
    x0, x1 = ranges[comm.rank]
    mesh = lib.disp_mesh(prm.disp_hw, fix_s.device)
    L = (2 * prm.disp_hw + 1) ** 3
    ssd = torch.empty((L, x1 - x0, w, d), dtype=torch.float32, device=fix_s.device)      # reused for the backward direction

    # The slab path takes the exact cost volume unless the certified mode is asked for explicitly (Params.mode): with
    # the host-run selection loop the certified pieces measured slower here (34.8 vs 31.0 ms for the config-2 pair on two
    # GPUs); they exist, and are parity-tested, for slabs whose exact volume would be the bottleneck.
    certified = prm.mode == "certified"

    def one_direction(f, m):
        lib.correlate_slab(f, m, prm.disp_hw, x0, x1, want_argmin=False, out=ssd, fast=certified)
        soft_slab = coupled_convex_slab(lib, comm, ssd, mesh, prm.disp_hw, (f, m, x0) if certified else None)
        return comm.gather_planes(soft_slab, ranges)

    soft = one_direction(fix_s, mov_s)
    if prm.ic:
        soft_ = one_direction(mov_s, fix_s)
        disp = pipeline.ic_and_upsample(lib, soft, soft_, shape, prm)
    else:
        disp = soft
    del ssd
    if prm.lambda_weight > 0:
        if prm.selected_niter < 1:
            raise ValueError("selected_niter must be >= 1 when lambda_weight > 0")
        p_full = pipeline.adam_start(lib, disp, fix_a.shape[1:], prm)
        u = adam_slab(lib, comm, p_full, fix_a, mov_a, prm)
        disp = pipeline.adam_finish(lib, u, shape, prm)
    return disp
