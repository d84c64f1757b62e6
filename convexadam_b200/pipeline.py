"""Device-resident registration pipeline: the fused form of convex_adam_pt (reference
src/convexAdam/convex_adam_MIND.py:102-193) over libconvexadam_b200's kernels.

Everything between the host images and the dense field stays on the GPU, on the caller's current stream, with
no host synchronisation (the reference syncs at utils.py:61 and round-trips the field through the CPU at
MIND.py:156).  The cost volume (L,h,w,d) is allocated once and reused for the backward direction.
"""
import os
import threading
from dataclasses import dataclass

import torch

from ._native import Lib, get_lib


TIMER = None   # bench.py installs an object with .mark(name) to time stages with CUDA events on the current stream


def _mark(name):
    if TIMER is not None:
        TIMER.mark(name)


@dataclass
class Params:
    mind_r: int = 1
    mind_d: int = 2
    lambda_weight: float = 1.25
    grid_sp: int = 6
    disp_hw: int = 4
    selected_niter: int = 80
    selected_smooth: int = 0
    grid_sp_adam: int = 2
    ic: bool = True
    # "certified": approximate cost volume (ca_correlate_fast) + selections proven equal to the exact path's
    # (ca_coupled_convex_certified); "exact": the bit-exact cost volume (ca_correlate + ca_coupled_convex).  Both
    # give bit-identical argmin / selections / disp_soft; CONVEXADAM_B200_MODE overrides the default.
    mode: str = ""
    # The forward and the backward direction of the convex stage are independent until inverse consistency.  "1": issue them
    # on two CUDA streams (a second cost volume, 8.6 GB at config 2): the latency-bound kernels of one direction share the SMs
    # with the bandwidth-bound ones of the other, 24.1 instead of 25.9 ms for the stage (+5 % pairs/s), same bits.  "0": one
    # after the other on the caller's stream (what bench.py times stage by stage -- per-kernel timings of concurrently running
    # kernels mean nothing).  "" = CONVEXADAM_B200_OVERLAP if set, else automatic: two streams whenever the second volume can
    # be allocated.
    overlap: str = ""

    def resolved_overlap(self):
        """True / False / None (automatic)."""
        s = self.overlap or os.environ.get("CONVEXADAM_B200_OVERLAP", "")
        if s in ("1", "on", "true", "True"):
            return True
        if s in ("0", "off", "false", "False"):
            return False
        if s in ("", "auto"):
            return None
        raise ValueError(f"unknown overlap setting {s!r} (0 | 1 | auto)")

    def resolved_mode(self):
        m = self.mode or os.environ.get("CONVEXADAM_B200_MODE", "certified")
        if m not in ("certified", "exact"):
            raise ValueError(f"unknown mode {m!r} (certified | exact)")
        return m


def convex_stage(lib: Lib, fix_s, mov_s, shape, prm: Params, ssd_buf=None):
    """correlate + coupled_convex (+ inverse consistency + upsampling when prm.ic).  fix_s/mov_s: (C,h,w,d).
    Returns disp_hr (3,H,W,D) in voxels (or the coarse (3,h,w,d) field in coarse units when ic=False -- the
    reference's behaviour at MIND.py:143-144)."""
    H, W, D = shape
    C, h, w, d = fix_s.shape
    dev = fix_s.device
    mesh = lib.disp_mesh(prm.disp_hw, dev)
    L = (2 * prm.disp_hw + 1) ** 3
    if ssd_buf is None or ssd_buf.shape != (L, h, w, d):
        ssd_buf = torch.empty((L, h, w, d), dtype=torch.float32, device=dev)
    certified = prm.resolved_mode() == "certified"

    def one_direction(a, b, buf, mark=True):
        # the argmin of utils.py:87 is taken inside coupled_convex, in the pass that builds its block minima
        ssd, _ = lib.correlate(a, b, prm.disp_hw, want_argmin=False, out=buf, fast=certified)
        if mark:
            _mark("correlate")
        soft = lib.coupled_convex_certified(ssd, a, b, mesh, prm.disp_hw) if certified else lib.coupled_convex(ssd, None, mesh, prm.disp_hw)
        if mark:
            _mark("coupled_convex")
        return soft

    if not prm.ic:
        return one_direction(fix_s, mov_s, ssd_buf)
    want = prm.resolved_overlap()
    ssd_buf2 = None
    main = torch.cuda.current_stream(dev) if lib.is_cuda else None
    if lib.is_cuda and want is not False:
        try:
            ssd_buf2 = _second_volume(ssd_buf, main)
        except torch.cuda.OutOfMemoryError:
            if want:
                raise
    if ssd_buf2 is None:
        soft = one_direction(fix_s, mov_s, ssd_buf)
        soft_ = one_direction(mov_s, fix_s, ssd_buf)
        return ic_and_upsample(lib, soft, soft_, shape, prm)
    # Each direction owns a cost volume; both streams start after, and are joined by, the caller's stream.  The second volume
    # and the side streams belong to (device, caller's stream): calls on one stream reuse them in stream order; the lock makes
    # the enqueue of one call's fork .. join atomic, so that two host threads on the same stream cannot interleave them.
    side = _side_streams(dev)
    outs = []
    with _FORK_LOCK:
        start = torch.cuda.Event()
        start.record(main)
        for st, (a, b, buf) in zip(side, ((fix_s, mov_s, ssd_buf), (mov_s, fix_s, ssd_buf2))):
            st.wait_event(start)
            with torch.cuda.stream(st):
                outs.append(one_direction(a, b, buf, mark=False))
            done = torch.cuda.Event()
            done.record(st)
            main.wait_event(done)
    for t in (fix_s, mov_s, ssd_buf):
        for st in side:
            t.record_stream(st)
    for t in outs:                 # allocated on a side stream, consumed on the caller's
        t.record_stream(main)
    _mark("correlate+coupled_convex (two directions on two streams)")
    return ic_and_upsample(lib, outs[0], outs[1], shape, prm)


_SIDE = {}
_VOL2 = {}                          # (device, shape, caller's stream) -> the backward direction's cost volume
_FORK_LOCK = threading.Lock()


def _second_volume(buf, main):
    """The backward direction's cost volume, kept per (device, shape, caller's stream) -- 8.6 GB at config 2.  Allocating it per
    call would be free for a caller that synchronises between calls, but a caller that enqueues registrations ahead of the GPU
    would hold one volume per call in flight.  At most two are kept."""
    key = (str(buf.device), tuple(buf.shape), int(main.cuda_stream))
    vol = _VOL2.pop(key, None)
    if vol is None:
        while len(_VOL2) >= 2:
            _VOL2.pop(next(iter(_VOL2)))
        vol = torch.empty_like(buf)
    _VOL2[key] = vol                # most recently used last
    return vol


def release_buffers():
    """Drops the cached second cost volumes (they are re-allocated on the next call)."""
    _VOL2.clear()


def _side_streams(dev):
    key = str(dev)
    if key not in _SIDE:
        _SIDE[key] = (torch.cuda.Stream(dev), torch.cuda.Stream(dev))
    return _SIDE[key]


def ic_and_upsample(lib: Lib, soft, soft_, shape, prm: Params):
    """Inverse consistency of the two coarse fields and upsampling to full resolution (MIND.py:133-141)."""
    H, W, D = shape
    _, h, w, d = soft.shape
    # scale = tensor([h-1, w-1, d-1]) / 2 (MIND.py:134); (disp/scale).flip(1) (:139)
    scale = tuple(float(torch.tensor(float(s - 1)) / 2) for s in (h, w, d))
    a = lib.field_affine(soft, flip=True, div=scale)
    b = lib.field_affine(soft_, flip=True, div=scale)
    ice, _ = lib.inverse_consistency(a, b, 15)
    # disp_ice.flip(1) * scale * grid_sp, upsampled (:141); mul1 is indexed by the SOURCE channel
    up = lib.field_affine(ice, flip=True, mul1=scale[::-1], mul2=float(prm.grid_sp))
    out = lib.trilinear(up, (H, W, D))
    _mark("inverse_consistency+upsample")
    return out


def adam_stage(lib: Lib, disp_hr, fix_a, mov_a, shape, prm: Params):
    """Adam instance optimisation (MIND.py:147-191).  disp_hr: (3,*,*,*) start field; fix_a/mov_a: (C,ha,wa,da)."""
    H, W, D = shape
    if prm.selected_niter < 1:
        # the reference reads `disp_sample` after a loop that never ran (MIND.py:181) and dies with a NameError
        raise ValueError("selected_niter must be >= 1 when lambda_weight > 0")
    ha, wa, da = fix_a.shape[1:]
    gsa = float(prm.grid_sp_adam)
    p = adam_start(lib, disp_hr, (ha, wa, da), prm)
    m = torch.zeros_like(p)
    v = torch.zeros_like(p)
    u = lib.adam_run(fix_a, mov_a, float(prm.lambda_weight), p, m, v, 1, int(prm.selected_niter))
    return adam_finish(lib, u, shape, prm)


def adam_start(lib: Lib, disp_hr, grid, prm: Params):
    """Initial Adam parameters: the start field resized to the Adam grid, in Adam-grid voxels (MIND.py:153,156)."""
    gsa = float(prm.grid_sp_adam)
    return lib.field_affine(lib.trilinear(disp_hr, tuple(grid)), div=(gsa, gsa, gsa))


def adam_finish(lib: Lib, u, shape, prm: Params):
    """u of the last forward -> dense field in voxels, optional 3x box smoothing (MIND.py:181-191)."""
    H, W, D = shape
    gsa = float(prm.grid_sp_adam)
    out = lib.trilinear(lib.field_affine(u, mul2=gsa), (H, W, D))                           # :181-182
    if prm.selected_smooth > 0:
        if prm.selected_smooth % 2 == 0:
            print("selected_smooth should be an odd number, adding 1")   # and then, like the reference, not adding it (:185-189)
        k = int(prm.selected_smooth)
        for _ in range(3):
            out = lib.avgpool3d(out, k, 1, k // 2)
    _mark("adam")
    return out


def register_features(lib: Lib, fix_s, mov_s, fix_a, mov_a, shape, prm: Params, ssd_buf=None):
    disp = convex_stage(lib, fix_s, mov_s, shape, prm, ssd_buf)
    if prm.lambda_weight > 0:
        disp = adam_stage(lib, disp, fix_a, mov_a, shape, prm)
    return disp


def register_mind(fixed_dev, moving_dev, prm: Params, lib: Lib = None, ssd_buf=None, moving_ready=None):
    """fixed_dev, moving_dev: (H,W,D) float32 CUDA tensors.  Returns the dense field (3,H',W',D') float32 on the
    device (voxel units, channel c = displacement along array axis c).  moving_ready: optional CUDA event after which
    moving_dev holds its data (its host-to-device copy may still be in flight on another stream while the fixed image's
    descriptors are computed)."""
    lib = lib or get_lib()
    H, W, D = fixed_dev.shape
    _mark("start")
    same = prm.grid_sp == prm.grid_sp_adam
    need_adam = prm.lambda_weight > 0
    gs_b = 0 if (same or not need_adam) else prm.grid_sp_adam
    fix_s, fix_a = lib.mindssc_pooled(fixed_dev, prm.mind_r, prm.mind_d, prm.grid_sp, gs_b)
    if moving_ready is not None:
        torch.cuda.current_stream(fixed_dev.device).wait_event(moving_ready)
    mov_s, mov_a = lib.mindssc_pooled(moving_dev, prm.mind_r, prm.mind_d, prm.grid_sp, gs_b)
    if same:
        fix_a, mov_a = fix_s, mov_s
    _mark("mind")
    return register_features(lib, fix_s, mov_s, fix_a, mov_a, (H, W, D), prm, ssd_buf)
