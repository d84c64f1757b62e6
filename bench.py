#!/usr/bin/env python
"""bench.py -- volume-pairs/s of the convexAdam hot path on B200 (BASELINE.json metric, config 2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--niter 100] [--shape 192 160 256]

One step = one full registration of one synthetic 192x160x256 pair (MIND-SSC, disp_hw=6, grid_sp=2, grid_sp_adam=2,
100 Adam iterations, ic=True, fp32).  `value` times the device-resident path (inputs already in HBM, field left in
HBM); `e2e` times the public API convex_adam_pt with pinned HOST inputs and the (H,W,D,3) float64 result copied back.
N>1 (torchrun): pairs are independent, every rank registers its own pairs, no data-path collective ("weak" scaling).
`--impl reference`: the reference's OWN torch-CPU path (the unmodified package staged under oracle/_ref by
oracle/stage_reference.py, imported through oracle/ref_shim.py) on all host threads, on a bounded sample of the workload
(a pair with 1/8 or 1/64 of the voxels, same hyper-parameters; cost is linear in voxels), rank 0 only; the C port
(oracle/ca_oracle.c) is timed next to it as a second figure, and alone when the staged reference is absent.
Extra keys of our arm: `roofline` = the correlate stage against SURVEY 8(d)'s bytes_corr (north_star's quantity),
`other_configs` (BASELINE configs 1 and 4 and the package-defaults regime, N=1), `reference_torch_cuda` (the reference's
torch-CUDA path on the same GPU, fp32 and fp16: the honest bar of SURVEY 8(d)), `slab_split` (N>1: ONE pair split into H
slabs over the N ranks, BASELINE config 5's way, with a bit-equality flag against the unsplit field).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NORTH = dict(mind_r=1, mind_d=2, lambda_weight=1.25, grid_sp=2, disp_hw=6, selected_smooth=0, grid_sp_adam=2, ic=True)
METRIC = "volume-pairs/sec @192x160x256 MIND+convex+Adam"
REF_WHAT = "the reference's own convex_adam_pt (unmodified multimodallearning/convexAdam package, staged byte for byte under oracle/_ref; torch CPU fp32)"
WORKLOAD = "%dx%dx%d pair, MIND-SSC r1 d2, disp_hw=6, grid_sp=2, grid_sp_adam=2, %d Adam it., ic=True, fp32 cost volume"


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                       "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return None
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().strip().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 8:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if not sm:
            return None
        sm_sorted = sorted(sm)
        return {"sm_mhz": sm_sorted[len(sm_sorted) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


class StageTimer:
    def __init__(self, torch):
        self.torch = torch
        self.events = []

    def mark(self, name):
        e = self.torch.cuda.Event(enable_timing=True)
        e.record()
        self.events.append((name, e))

    def totals(self):
        out = {}
        for (n0, e0), (n1, e1) in zip(self.events[:-1], self.events[1:]):
            if n1 == "start":
                continue
            out[n1] = out.get(n1, 0.0) + e0.elapsed_time(e1)
        return out


def algorithmic_bytes(shape, prm, C=12):
    """SURVEY.md §8(d), fp32 (s_f = s_s = 4); per-kernel figures as stated in DESIGN.md §4."""
    H, W, D = shape
    gs, hw = prm["grid_sp"], prm["disp_hw"]
    n = (H // gs) * (W // gs) * (D // gs)
    L = (2 * hw + 1) ** 3
    corr_dir = 2 * C * n * 4 + L * n * 4 + 8 * n          # both feature volumes once, ssd once, int64 argmin
    cc_pass = L * n * 4 + 12 * n + 8 * n                  # ssd once, disp_soft, selection
    na = (H // prm["grid_sp_adam"]) * (W // prm["grid_sp_adam"]) * (D // prm["grid_sp_adam"])
    return dict(n=n, L=L, n_adam=na, correlate_per_direction=corr_dir, coupled_per_pass=cc_pass,
                box_per_direction=2 * L * n * 4,          # box filters: every raw value read once, every cost written once
                blockmin_per_direction=L * n * 4 + (2 * hw + 1) ** 2 * n * 4 + 8 * n)   # cost volume read once; block minima + argmin written


def sample_shape(shape, scale_div):
    return tuple(s // scale_div for s in shape)


def cpu_port_sample(shape, prm, niter, scale_div):
    """Times the C port of the reference (oracle/ca_oracle.c) on a pair with 1/scale_div^3 of the voxels (same
    hyper-parameters); returns (pairs_per_s scaled to the full shape, seconds of the sample)."""
    from convexadam_b200.synth import make_pair
    from oracle import ca_oracle as orc
    H, W, D = sample_shape(shape, scale_div)
    fixed, moving = make_pair((H, W, D), seed=0)
    t0 = time.perf_counter()
    orc.convex_adam_pt(fixed.numpy(), moving.numpy(), prm["mind_r"], prm["mind_d"], prm["lambda_weight"], prm["grid_sp"],
                       prm["disp_hw"], niter, prm["selected_smooth"], prm["grid_sp_adam"], prm["ic"])
    dt = time.perf_counter() - t0
    ratio = (shape[0] * shape[1] * shape[2]) / float(H * W * D)
    return 1.0 / (dt * ratio), dt


def cpu_reference_sample(shape, prm, niter, scale_div, ref_mind):
    """The same sample through the reference's own convex_adam_pt on the CPU (fp32, every host thread)."""
    import torch
    from convexadam_b200.synth import make_pair
    H, W, D = sample_shape(shape, scale_div)
    fixed, moving = make_pair((H, W, D), seed=0)
    t0 = time.perf_counter()
    ref_mind.convex_adam_pt(fixed, moving, selected_niter=niter, device=torch.device("cpu"), dtype=torch.float32, **prm)
    dt = time.perf_counter() - t0
    ratio = (shape[0] * shape[1] * shape[2]) / float(H * W * D)
    return 1.0 / (dt * ratio), dt


def sample_description(shape, scale_div, what):
    H, W, D = sample_shape(shape, scale_div)
    ratio = (shape[0] * shape[1] * shape[2]) / float(H * W * D)
    return (f"one {H}x{W}x{D} pair (1/{ratio:.0f} of the voxels, same disp_hw/grid_sp/iterations) through {what}; "
            f"pairs/s scaled by 1/{ratio:.0f} (every stage of the path is linear in voxels at fixed disp_hw)")


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm is meant to use every host core (set before the
    OpenMP runtime of the oracle / torch initialises)."""
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count())


def import_staged_reference():
    """(utils, MIND) modules of the unmodified reference, or None when neither /root/reference nor oracle/_ref exists."""
    try:
        from oracle import ref_shim
        if not ref_shim.available():
            return None
        return ref_shim.import_reference()
    except Exception as exc:     # a missing dependency of the reference, etc.: report, fall back to the port
        sys.stderr.write(f"[bench] reference not importable: {exc}\n")
        return None


def run_reference(args, shape, prm):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    use_all_host_threads()
    from oracle import ca_oracle as orc
    orc.build()
    ref = import_staged_reference()
    cores = os.cpu_count()
    if ref is not None:
        import torch
        torch.set_num_threads(cores)
    total = args.steps + args.warmup
    # bounded sample: ~60 s (1/8 of the voxels) per step for short runs, ~8 s (1/64) otherwise, so that the whole run
    # ends within a few minutes whatever --steps / --warmup the driver passes
    scale_div = 2 if total <= 3 else 4
    vals, secs = [], []
    for i in range(total):
        v, dt = cpu_reference_sample(shape, prm, args.niter, scale_div, ref[1]) if ref is not None else \
            cpu_port_sample(shape, prm, args.niter, scale_div)
        if i >= args.warmup:
            vals.append(v); secs.append(dt)
    ratio = 1.0 / (vals[0] * secs[0])
    t = sum(secs) / len(secs)
    value = 1.0 / (t * ratio)
    kind = "reference" if ref is not None else "port"
    what = (REF_WHAT if ref is not None else "the C port of the reference (oracle/ca_oracle.c, OpenMP)")
    cpu = {"value": value, "unit": "pairs/s", "cores": cores, "kind": kind, "sample": sample_description(shape, scale_div, what),
           "sample_seconds": t}
    if ref is not None:      # the port beside it, once
        pv, pdt = cpu_port_sample(shape, prm, args.niter, scale_div)
        cpu["port"] = {"value": pv, "unit": "pairs/s", "sample_seconds": pdt,
                       "note": "oracle/ca_oracle.c (C restatement pinned to the reference's golden vectors) on the same sample"}
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t * ratio, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD % (tuple(shape) + (args.niter,)),
                   "note": "CPU arm: " + what + " on all host threads; each step is the bounded sample described in cpu_baseline"},
        "cpu_baseline": cpu,
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=("ours", "reference"))
    ap.add_argument("--niter", type=int, default=100)
    ap.add_argument("--shape", type=int, nargs=3, default=[192, 160, 256])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip other_configs / reference_torch_cuda / slab_split")
    args = ap.parse_args()
    shape = tuple(args.shape)
    prm = dict(NORTH)

    if args.impl == "reference":
        run_reference(args, shape, prm)
        return

    import torch
    import torch.distributed as dist

    from convexadam_b200 import pipeline
    from convexadam_b200._native import get_lib
    from convexadam_b200.convex_adam_MIND import convex_adam_pt
    from convexadam_b200.synth import make_pair

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    lib = get_lib()
    steps, warmup = args.steps, max(args.warmup, 3)

    fixed, moving = make_pair(shape, seed=rank)
    fixed_pin, moving_pin = fixed.pin_memory(), moving.pin_memory()
    fixed_dev, moving_dev = fixed_pin.to(dev), moving_pin.to(dev)
    # `value` and the stage / kernel timings: the two directions of the convex stage one after the other on one stream
    # (overlap="0"); the library's default issues them on two streams (Params.overlap), which is what `e2e` and
    # `value_two_stream_directions` measure -- per-kernel timings of concurrently running kernels mean nothing
    P = pipeline.Params(selected_niter=args.niter, overlap="0", **prm)
    ab = algorithmic_bytes(shape, prm)
    ssd_buf = torch.empty((ab["L"],) + tuple(s // prm["grid_sp"] for s in shape), dtype=torch.float32, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput (`value`) ----
    for _ in range(warmup):
        pipeline.register_mind(fixed_dev, moving_dev, P, lib, ssd_buf)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    timer = StageTimer(torch)
    pipeline.TIMER = timer
    n0 = lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(steps):
        field = pipeline.register_mind(fixed_dev, moving_dev, P, lib, ssd_buf)
    e1.record()
    barrier()
    pipeline.TIMER = None
    launches = lib.launch_count() - n0
    ms = max_over_ranks(e0.elapsed_time(e1))
    # Per-kernel times (`roofline_other_kernels`): a second, instrumented pass -- the library brackets the launches of its main
    # kernels with CUDA events on its launch stream.  Kept out of the timed region above: an event between two kernels of the
    # Adam loop undoes their programmatic dependent launch (200 events per registration cost 1.4 ms of its 31 ms).
    ksteps = min(steps, 3)
    lib.kernel_timing(("raw_ssd", "box_filters", "corr_fused", "corr_pipe", "blockmin_argmin", "select_pass", "exact_redecision", "adam_grad"))
    ek0, ek1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ek0.record()
    for _ in range(ksteps):
        field = pipeline.register_mind(fixed_dev, moving_dev, P, lib, ssd_buf)
    ek1.record()
    barrier()
    ktimes = lib.kernel_times()
    lib.kernel_timing(())
    kstep_ms = ek0.elapsed_time(ek1) / ksteps
    clocks = sampler.stop() if sampler else None
    stages = {k: v / steps for k, v in timer.totals().items()}
    value = world * steps / (ms / 1e3)

    # ---- the reference's own timed region t0..t1: the same, plus the host-to-device copy of both images ----
    for _ in range(2):
        fixed_dev.copy_(fixed_pin, non_blocking=True); moving_dev.copy_(moving_pin, non_blocking=True)
        pipeline.register_mind(fixed_dev, moving_dev, P, lib, ssd_buf)
    barrier()
    e0.record()
    for _ in range(steps):
        fixed_dev.copy_(fixed_pin, non_blocking=True); moving_dev.copy_(moving_pin, non_blocking=True)
        field = pipeline.register_mind(fixed_dev, moving_dev, P, lib, ssd_buf)
    e1.record()
    barrier()
    value_h2d = {"value": world * steps / (max_over_ranks(e0.elapsed_time(e1)) / 1e3), "unit": "pairs/s",
                 "h2d_bytes_per_step": int(2 * fixed.numel() * 4)}

    # ---- the library's default mode: the two directions of the convex stage on two CUDA streams (a second cost volume) ----
    value_overlap = None
    if not args.no_extras:
        P2 = pipeline.Params(selected_niter=args.niter, overlap="1", **prm)
        for _ in range(2):
            pipeline.register_mind(fixed_dev, moving_dev, P2, lib, ssd_buf)
        barrier()
        e0.record()
        for _ in range(steps):
            field = pipeline.register_mind(fixed_dev, moving_dev, P2, lib, ssd_buf)
        e1.record()
        barrier()
        value_overlap = {"value": world * steps / (max_over_ranks(e0.elapsed_time(e1)) / 1e3), "unit": "pairs/s",
                         "note": "forward and backward direction of the convex stage issued on two streams (the library's default, "
                                 "what convex_adam_pt and hence `e2e` run); `value` is the one-stream mode the stage timings need"}
        torch.cuda.empty_cache()

    # ---- end to end through the public API (host buffers in, host array out) ----
    e2e = None
    if not args.no_e2e:
        kw = dict(prm, selected_niter=args.niter, dtype=torch.float32, device=dev)
        for _ in range(3):      # steady state: the caller still holds the previous result when the next call starts
            out = convex_adam_pt(fixed_pin, moving_pin, **kw)
        barrier()
        t0 = time.perf_counter()
        # N > 1: the ranks are independent workers, not a lockstep job.  Started together they would all copy their 189 MB
        # result to the host at the same instants and queue on the host's memory path; each rank therefore starts rank/N of
        # a step late (at most one copy time per rank) -- INSIDE the timed region, so the delay is paid for in the number reported.
        stagger = 0.0
        if world > 1 and os.environ.get("BENCH_E2E_STAGGER", "1") == "1":
            # spacing = one rank's copy time at full PCIe rate (enough to take the copies apart), at most 1/N of a step
            copy_s = (2 * fixed.numel() * 4 + 3 * fixed.numel() * 8) / 50e9
            stagger = rank * min((ms / steps / 1e3) / world, copy_s)
            time.sleep(stagger)
        for _ in range(steps):
            out = convex_adam_pt(fixed_pin, moving_pin, **kw)
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": world * steps / dt, "unit": "pairs/s", "h2d_bytes_per_step": int(2 * fixed.numel() * 4),
               "d2h_bytes_per_step": int(out.nbytes), "ms_per_step": 1e3 * dt / steps}
        if world > 1:
            e2e["note"] = ("ranks start rank/N of a step apart inside the timed region (independent workers; in lockstep their "
                           "device-to-host copies collide on the host's memory path)" if stagger or rank == 0 and os.environ.get("BENCH_E2E_STAGGER", "1") == "1"
                           else "ranks start in lockstep")

    # ---- the same end-to-end path with two pairs in flight (batch API, config 3's way of running many pairs) ----
    piped = None
    if not args.no_e2e:
        from convexadam_b200.batch import iter_register_pipelined
        kwp = dict(prm, selected_niter=args.niter)
        checksum = 0.0
        for fld in iter_register_pipelined([(fixed_pin, moving_pin)] * 4, depth=2, device=dev, **kwp):   # warm-up: per-stream pools, pinned buffers
            checksum += float(fld[0, 0, 0, 0])
        barrier()
        npipe = max(6, steps)
        t0 = time.perf_counter()
        for fld in iter_register_pipelined([(fixed_pin, moving_pin)] * npipe, depth=2, device=dev, **kwp):
            checksum += float(fld[0, 0, 0, 0])       # the consumer touches each field, then drops it
        torch.cuda.synchronize()
        dtp = max_over_ranks(time.perf_counter() - t0)
        piped = {"value": world * npipe / dtp, "unit": "pairs/s", "pairs": npipe, "in_flight": 2, "ms_per_pair": 1e3 * dtp / npipe,
                 "note": "batch.iter_register_pipelined: host images in, host fields out, two pairs in flight on two streams per GPU"}

    # ---- N>1: ONE pair split into H slabs over the ranks (BASELINE config 5's way; every rank takes part) ----
    slab_split = None
    if world > 1 and not args.no_extras:
        slab_split = measure_slab_split(torch, dist, lib, dev, world, rank, args.niter)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = peaks()
    ndir = 2 if prm["ic"] else 1
    mode = P.resolved_mode()
    step_ms = ms / steps

    def kernel_line(slot, name, bytes_per_step, bound, note):
        tot, cnt = ktimes.get(slot, (0.0, 0))
        if cnt == 0:
            return None
        per_step = tot / ksteps
        ach = bytes_per_step / (per_step * 1e-3) / 1e9
        return {"kernel": name, "bound": bound, "achieved": ach, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                "frac": ach / peak, "launches_per_step": cnt / ksteps, "avg_launch_ms": tot / cnt,
                "algorithmic_bytes_per_launch": bytes_per_step / (cnt / ksteps), "ms_per_step": per_step,
                "share_of_step": per_step / kstep_ms, "note": note + " [instrumented pass: %.2f ms per step]" % kstep_ms}

    # Headline roofline = north_star's quantity: the correlation stage of one direction (every kernel `correlate`
    # launches, timed with CUDA events on the launch stream inside the timed region) against SURVEY 8(d)'s bytes_corr.
    t_corr = stages.get("correlate", 0.0) / ndir
    ach_corr = ab["correlate_per_direction"] / (t_corr * 1e-3) / 1e9 if t_corr > 0 else 0.0
    fused = "corr_fused" in ktimes
    pipe_kernel = "corr_pipe" in ktimes
    corr_kernels = "k_corr_fused (one fused launch)" if fused else "k_corr_pipe (one launch of producer / consumer CTAs)" if pipe_kernel else \
        ("k_raw_ssd_tmap + k_box12_fast" if mode == "certified" else "k_raw_ssd_staged + k_box12") + ", %d launches" % int(
            round((ktimes.get("raw_ssd", (0, 0))[1] + ktimes.get("box_filters", (0, 0))[1]) / ksteps / ndir))
    roof = {"kernel": "correlate stage of one direction (%s): features in, (L,h,w,d) cost volume out" % corr_kernels,
            "bound": "hbm", "achieved": ach_corr, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
            "frac": ach_corr / peak if peak else None, "launches_per_step": ndir, "avg_launch_ms": t_corr,
            "algorithmic_bytes_per_launch": ab["correlate_per_direction"], "ms_per_step": ndir * t_corr,
            "share_of_step": ndir * t_corr / step_ms,
            "note": "algorithmic bytes = SURVEY 8(d) bytes_corr per direction: both coarse feature volumes read once, the cost volume "
                    "written once (+ int64 argmin); 'launch' = one direction of the stage",
            "traffic": committed_traffic("pipe" if pipe_kernel else mode, fused, shape)}
    # per-kernel extras, each against the SURVEY 8(d) bytes it is responsible for (the raw-SSD planes between the two kernels
    # of the stage are NOT algorithmic bytes: they are the traffic the fused kernel removes)
    others = [kernel_line("corr_fused", "k_corr_fused (raw SSD + both box filters + H march, tensor-map TMA)",
                          ndir * ab["correlate_per_direction"], "hbm", "writes the cost volume once; features are L2 hits"),
              kernel_line("corr_pipe", "k_corr_pipe (producer / consumer CTAs, raw SSD planes through a ring in L2)",
                          ndir * ab["correlate_per_direction"], "hbm", "writes the cost volume once; the raw planes stay in L2"),
              kernel_line("box_filters", "k_box12_fast (both 3x3x3 box filters, separable)" if mode == "certified" else
                          "k_box12 (both 3x3x3 box filters, reference summation order)", ndir * (ab["L"] * ab["n"] * 4), "hbm",
                          "bytes = the cost volume it writes; it also reads the raw-SSD planes (same size, not algorithmic)"),
              kernel_line("raw_ssd", "k_raw_ssd_tmap (SSD over 13 kD x 4 kW per CTA; tiles staged by tensor-map TMA from the unpadded features)" if mode == "certified"
                          else "k_raw_ssd_staged (SSD over 13 kD x 4 kW per CTA from bulk-copy staged tiles)",
                          ndir * (2 * 12 * ab["n"] * 4), "hbm",
                          "bytes = the two feature volumes it reads; it writes the raw-SSD planes (not algorithmic); bound by fp32 FMA lanes: 24 lane-ops per entry"),
              kernel_line("blockmin_argmin", "k_blockmin_argmin (argmin + block/slab minima, one streaming read of the volume)",
                          ndir * ab["blockmin_per_direction"], "hbm", "streaming pass"),
              kernel_line("select_pass", "k_cost_argmin_sep (one coupled-convex selection pass, exact pruning, separable grid)",
                          ndir * 6 * ab["coupled_per_pass"], "hbm",
                          "bytes = what an unpruned pass reads (SURVEY 8d); the pruned pass touches ~5 % of them, hence frac > 1"),
              kernel_line("exact_redecision", "k_cc_fixup (exact re-decision of near-ties)", 0, "hbm", "latency-bound, a few hundred voxels"),
              kernel_line("adam_grad", "k_adam_grad (trilinear sample + SSD residual + analytic gradient + regulariser)",
                          args.niter * ab["n_adam"] * (2 * 12 * 4 + 24), "hbm", "bytes = both feature volumes + u in, g_u out, per iteration")]
    line = {
        "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD % (shape + (args.niter,)),
                   "mode": mode + (" (approximate cost volume, selections certified bit-equal to the exact path)" if mode == "certified" else ""),
                   "pairs_per_step_per_gpu": 1, "parallelism": f"pair-parallel x{world} (no collective on the data path)",
                   "value_region": "inputs resident in HBM when the timed region starts (bench contract); `value_t0_t1` adds the "
                                   "host-to-device copy of both images, i.e. the reference's own t0..t1 (MIND.py:102-193)",
                   "streams": "`value`, `stages_ms`, `roofline`: the two directions of the convex stage one after the other on one stream, "
                              "so that every stage and kernel is timed alone; the library's default (what `e2e` runs) issues them on "
                              "two streams: `value_two_stream_directions`",
                   "l2": "the 8.6 GB cost volume (>> 126 MB L2) is rewritten and re-read every step; no explicit flush"},
        "stages_ms": stages,
        "value_t0_t1": value_h2d,
        "value_two_stream_directions": value_overlap,
        "roofline": roof,
        "roofline_other_kernels": [k for k in others if k],
        "gpu_launches": int(launches), "clocks": clocks, "e2e": e2e, "e2e_pipelined": piped,
    }
    if slab_split is not None:
        line["slab_split"] = slab_split
    if world == 1 and not args.no_extras:
        del ssd_buf, field
        torch.cuda.empty_cache()
        line["other_configs"] = measure_other_configs(torch, dev)
        line["reference_torch_cuda"] = measure_reference_cuda(torch, dev, shape, prm, args.niter)
    if not args.no_cpu_baseline and world == 1:
        import ctypes
        try:    # torch's OpenMP runtime is already up: raise its thread count directly (the oracle shares libgomp)
            ctypes.CDLL("libgomp.so.1").omp_set_num_threads(os.cpu_count())
        except OSError:
            pass
        torch.set_num_threads(os.cpu_count())
        from oracle import ca_oracle as orc
        orc.build()
        ref = import_staged_reference()
        if ref is not None:
            v, dt = cpu_reference_sample(shape, prm, args.niter, 4, ref[1])
            what, kind = REF_WHAT, "reference"
        else:
            v, dt = cpu_port_sample(shape, prm, args.niter, 2)
            what, kind = "the C port of the reference (oracle/ca_oracle.c, OpenMP)", "port"
        line["cpu_baseline"] = {"value": v, "unit": "pairs/s", "cores": os.cpu_count(), "kind": kind,
                                "sample": sample_description(shape, 4 if ref is not None else 2, what), "sample_seconds": dt}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def committed_traffic(mode, fused, shape):
    """DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of the correlate stage per direction from the committed
    `ncu --set full` captures of this commit's kernels at this shape (profiles/r2_traffic.json), else None."""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_traffic.json")) as f:
            t = json.load(f)
        if tuple(t["shape"]) != tuple(shape):
            return None
        return float(t["fused" if fused else mode]["dram_bytes_per_direction"])
    except Exception:
        return None


def time_calls(torch, fn, reps):
    for _ in range(3):      # steady state: workspaces, pinned result buffers and the caching allocator have settled
        out = fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0) / reps, out


def measure_other_configs(torch, dev):
    """BASELINE configs 1 and 4 and the package-defaults regime through the public entry points (host arrays in, host field
    out; wall clock around synchronised calls, 3 warm-ups)."""
    from convexadam_b200 import convex_adam_nnUNet
    from convexadam_b200.convex_adam_MIND import convex_adam_pt
    from convexadam_b200.synth import make_label_pair, make_pair
    out = {}
    ns = dict(NORTH)
    dflt = dict(mind_r=1, mind_d=2, lambda_weight=1.25, grid_sp=6, disp_hw=4, selected_smooth=0, grid_sp_adam=2, ic=True)
    f64, m64 = make_pair((64, 64, 64), seed=21, disp_scale=0.08)
    ms_, _ = time_calls(torch, lambda: convex_adam_pt(f64, m64, selected_niter=100, dtype=torch.float32, device=dev, **ns), 5)
    out["config1_64cubed_north_star_params"] = {"ms_per_pair_e2e": ms_, "params": "grid_sp=2 disp_hw=6 100 it."}
    ms_, _ = time_calls(torch, lambda: convex_adam_pt(f64, m64, selected_niter=80, dtype=torch.float32, device=dev, **dflt), 5)
    out["config1_64cubed_package_defaults"] = {"ms_per_pair_e2e": ms_, "params": "grid_sp=6 disp_hw=4 80 it."}
    f, m = make_pair((192, 160, 256), seed=0)
    f, m = f.pin_memory(), m.pin_memory()
    ms_, _ = time_calls(torch, lambda: convex_adam_pt(f, m, selected_niter=80, dtype=torch.float32, device=dev, **dflt), 3)
    out["defaults_regime_192x160x256"] = {"ms_per_pair_e2e": ms_, "params": "grid_sp=6 disp_hw=4 80 it. (Adam-dominated regime, SURVEY 8d)"}
    del f, m
    lf, lm = make_label_pair((192, 192, 192), n_labels=16, seed=0)
    ms_, fld = time_calls(torch, lambda: convex_adam_nnUNet.convex_adam_pt(lf, lm, 1.25, 2, 6, 100, 0, 2, True, dtype=torch.float32, device=dev), 2)
    out["config4_nnunet_192cubed_16_labels"] = {"ms_per_pair_e2e": ms_, "params": "C=16 one-hot, grid_sp=2 disp_hw=6 100 it., exact cost volume",
                                                "field_absmax": float(abs(fld).max())}
    torch.cuda.empty_cache()
    return out


def measure_reference_cuda(torch, dev, shape, prm, niter):
    """The reference's own torch-CUDA path (unmodified package from oracle/_ref) on this GPU at the bench workload, fp32 and
    its default fp16: SURVEY 8(d)'s 'honest bar'.  Informational; absent reference -> {"unavailable": ...}."""
    ref = import_staged_reference()
    if ref is None:
        return {"unavailable": "oracle/_ref not staged (run oracle/stage_reference.py where /root/reference exists)"}
    from convexadam_b200.synth import make_pair
    fixed, moving = make_pair(shape, seed=0)
    out = {"note": "unmodified reference, torch %s CUDA path, same synthetic pair; wall clock around a synchronised call, 1 warm-up"
                   % torch.__version__}
    for name, dt in (("fp32", torch.float32), ("fp16_default", torch.float16)):
        try:
            torch.cuda.empty_cache()
            torch.cuda.reset_peak_memory_stats(dev)
            def call():
                return ref[1].convex_adam_pt(fixed, moving, selected_niter=niter, device=dev, dtype=dt, **prm)
            call()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            call()
            torch.cuda.synchronize()
            sec = time.perf_counter() - t0
            out[name] = {"ms_per_pair": 1e3 * sec, "pairs_per_s": 1.0 / sec, "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 2 ** 30}
        except Exception as exc:
            out[name] = {"failed": f"{type(exc).__name__}: {str(exc)[:200]}"}
    torch.cuda.empty_cache()
    return out


def measure_slab_split(torch, dist, lib, dev, world, rank, niter):
    """One pair split into H slabs over all ranks (convexadam_b200/slab.py): the config-2 pair, with a bit-equality flag
    against the unsplit single-GPU field; at 8 ranks also BASELINE config 5 (384^3, disp_hw=8), which fits no single GPU."""
    from convexadam_b200 import pipeline, slab
    from convexadam_b200.synth import make_pair

    def run(shape, hw, reps, compare):
        f, m = make_pair(shape, seed=0)
        fd, md = f.to(dev), m.to(dev)
        P = pipeline.Params(selected_niter=niter, **dict(NORTH, disp_hw=hw))
        out = slab.register_mind_slab(fd, md, P, lib)
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            out = slab.register_mind_slab(fd, md, P, lib)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res = {"shape": list(shape), "disp_hw": hw, "niter": niter, "ms_per_pair": float(t.item()), "scaling": "strong",
               "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 2 ** 30}
        if compare:
            same = torch.ones(1, dtype=torch.int32, device=dev)
            if rank == 0:
                ref = pipeline.register_mind(fd, md, P, lib)
                same[0] = int(bool((ref.view(torch.int32) == out.view(torch.int32)).all()))
                del ref
            dist.broadcast(same, 0)
            res["bit_equal_to_unsplit_field"] = bool(same.item())
        del out, fd, md
        torch.cuda.empty_cache()
        return res

    out = {"config2_pair": run((192, 160, 256), 6, 2, True)}
    if world >= 8:
        out["config5_384cubed_hw8"] = run((384, 384, 384), 8, 1, False)
        out["config5_384cubed_hw8"]["note"] = "cost volume 139 GB per direction: no unsplit run exists to compare with; parity of the slab pieces is tests/test_slab_*"
    return out


if __name__ == "__main__":
    main()
