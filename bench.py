#!/usr/bin/env python
"""bench.py -- volume-pairs/s of the convexAdam hot path on B200 (BASELINE.json metric, config 2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--niter 100] [--shape 192 160 256]

One step = one full registration of one synthetic 192x160x256 pair (MIND-SSC, disp_hw=6, grid_sp=2, grid_sp_adam=2,
100 Adam iterations, ic=True, fp32).  `value` times the device-resident path (inputs already in HBM, field left in
HBM); `e2e` times the public API convex_adam_pt with pinned HOST inputs and the (H,W,D,3) float64 result copied back.
N>1 (torchrun): pairs are independent, every rank registers its own pairs, no data-path collective ("weak" scaling).
`--impl reference`: the CPU oracle (oracle/ca_oracle.c, the restatement of the reference's torch-CPU path; the
reference itself is pure Python over torch and cannot travel to the GPU box) on a bounded sample, rank 0 only.
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
    return dict(n=n, L=L, correlate_per_direction=corr_dir, coupled_per_pass=cc_pass,
                box_per_direction=2 * L * n * 4,          # box filters: every raw value read once, every cost written once
                blockmin_per_direction=L * n * 4 + (2 * hw + 1) ** 2 * n * 4 + 8 * n)   # cost volume read once; block minima + argmin written


def cpu_reference_sample(shape, prm, niter, scale_div):
    """Times the CPU oracle on a volume with 1/scale_div^3 of the voxels (same hyper-parameters); returns
    (pairs_per_s_extrapolated, seconds, sample_description, threads)."""
    from convexadam_b200.synth import make_pair
    from oracle import ca_oracle as orc
    H, W, D = (s // scale_div for s in shape)
    fixed, moving = make_pair((H, W, D), seed=0)
    t0 = time.perf_counter()
    orc.convex_adam_pt(fixed.numpy(), moving.numpy(), prm["mind_r"], prm["mind_d"], prm["lambda_weight"], prm["grid_sp"],
                       prm["disp_hw"], niter, prm["selected_smooth"], prm["grid_sp_adam"], prm["ic"])
    dt = time.perf_counter() - t0
    ratio = (shape[0] * shape[1] * shape[2]) / float(H * W * D)
    desc = (f"one {H}x{W}x{D} pair (1/{ratio:.0f} of the voxels, same disp_hw/grid_sp/iterations) through the C oracle "
            f"(OpenMP); pairs/s scaled by 1/{ratio:.0f} (cost is linear in voxels)")
    return 1.0 / (dt * ratio), dt, desc, os.cpu_count()


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm is meant to use every host core (set before the
    OpenMP runtime of the oracle / torch initialises)."""
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count())


def run_reference(args, shape, prm):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    use_all_host_threads()
    from oracle import ca_oracle as orc
    orc.build()
    total = args.steps + args.warmup
    scale_div = 2 if total <= 3 else 4
    vals, secs, desc, cores = [], [], "", os.cpu_count()
    for i in range(total):
        v, dt, desc, cores = cpu_reference_sample(shape, prm, args.niter, scale_div)
        if i >= args.warmup:
            vals.append(v); secs.append(dt)
    t = sum(secs) / len(secs)
    ratio = 1.0 / (vals[0] * secs[0]) if vals else 1.0
    value = 1.0 / (t * ratio)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t * ratio, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%dx%dx%d pair, MIND-SSC r1 d2, disp_hw=6, grid_sp=2, grid_sp_adam=2, %d Adam it., ic=True, fp32 cost volume" % (tuple(shape) + (args.niter,)),
                   "note": "CPU arm: the reference is pure Python over torch and absent on the GPU box; this is its C restatement "
                           "(oracle/ca_oracle.c, pinned bit-exact to reference golden vectors) on all host threads"},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "kind": "port", "sample": desc},
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
    P = pipeline.Params(selected_niter=args.niter, **prm)
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
    # CUDA events around the launches of the convex stage's main kernels, recorded by the library on its launch stream
    lib.kernel_timing(("raw_ssd", "box_filters", "blockmin_argmin", "select_pass", "exact_redecision"))
    n0 = lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(steps):
        field = pipeline.register_mind(fixed_dev, moving_dev, P, lib, ssd_buf)
    e1.record()
    barrier()
    pipeline.TIMER = None
    ktimes = lib.kernel_times()
    lib.kernel_timing(())
    launches = lib.launch_count() - n0
    ms = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if sampler else None
    stages = {k: v / steps for k, v in timer.totals().items()}
    value = world * steps / (ms / 1e3)

    # ---- end to end through the public API (host buffers in, host array out) ----
    e2e = None
    if not args.no_e2e:
        kw = dict(prm, selected_niter=args.niter, dtype=torch.float32, device=dev)
        for _ in range(3):      # steady state: the caller still holds the previous result when the next call starts
            out = convex_adam_pt(fixed_pin, moving_pin, **kw)
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            out = convex_adam_pt(fixed_pin, moving_pin, **kw)
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": world * steps / dt, "unit": "pairs/s", "h2d_bytes_per_step": int(2 * fixed.numel() * 4),
               "d2h_bytes_per_step": int(out.nbytes), "ms_per_step": 1e3 * dt / steps}

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
        per_step = tot / steps
        ach = bytes_per_step / (per_step * 1e-3) / 1e9
        return {"kernel": name, "bound": bound, "achieved": ach, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                "frac": ach / peak, "launches_per_step": cnt / steps, "avg_launch_ms": tot / cnt,
                "algorithmic_bytes_per_launch": bytes_per_step / (cnt / steps), "ms_per_step": per_step,
                "share_of_step": per_step / step_ms, "note": note}

    box_name = "k_box12_fast (both 3x3x3 box filters of the cost volume, separable)" if mode == "certified" else \
               "k_box12 (both 3x3x3 box filters of the cost volume, reference summation order)"
    roof = kernel_line("box_filters", box_name, ndir * ab["box_per_direction"], "hbm",
                       "dominant kernel of the step; reads the raw SSD planes once, writes the cost volume once")
    # DRAM bytes per launch of the dominant kernel: the committed ncu capture of the same kernel at this shape
    # (profiles/r1_ncu_k_box12_fast.txt, a 52-pair launch: dram__bytes_read.sum + dram__bytes_write.sum = 5.580 GB against
    # 5.320 GB algorithmic, the difference being the zero halo rows it also reads) scaled to the average launch of this run
    if roof is not None:
        roof["traffic"] = (5.580 / 5.320) * roof["algorithmic_bytes_per_launch"] if (mode == "certified" and tuple(shape) == (192, 160, 256)) else None
    others = [kernel_line("raw_ssd", "k_raw_ssd_staged (SSD over 13 kD x 4 kW per CTA from TMA-staged tiles)",
                          ndir * (ab["L"] * ab["n"] * 4), "hbm",
                          "fp32-issue-bound by design: 24 flop per written float; bytes = the raw SSD planes it writes"),
              kernel_line("blockmin_argmin", "k_blockmin_argmin (argmin + block/slab minima, one streaming read of the volume)",
                          ndir * ab["blockmin_per_direction"], "hbm", "streaming pass"),
              kernel_line("select_pass", "k_cost_argmin_sep (one coupled-convex selection pass, exact pruning, separable grid)",
                          ndir * 6 * ab["coupled_per_pass"], "hbm",
                          "bytes = what an unpruned pass reads (SURVEY 8d); the pruned pass touches ~5 % of them, hence frac > 1"),
              kernel_line("exact_redecision", "k_cc_fixup (exact re-decision of near-ties)", 0, "hbm", "latency-bound, a few hundred voxels")]
    t_corr = stages.get("correlate", 0.0) / ndir
    ach_corr = ab["correlate_per_direction"] / (t_corr * 1e-3) / 1e9 if t_corr > 0 else 0.0
    line = {
        "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "%dx%dx%d pair, MIND-SSC r1 d2, disp_hw=6, grid_sp=2, grid_sp_adam=2, %d Adam it., ic=True, fp32 cost volume"
                               % (shape + (args.niter,)),
                   "mode": mode + (" (approximate cost volume, selections certified bit-equal to the exact path)" if mode == "certified" else ""),
                   "pairs_per_step_per_gpu": 1, "parallelism": f"pair-parallel x{world} (no collective on the data path)",
                   "l2": "the 8.6 GB cost volume (>> 126 MB L2) is rewritten and re-read every step; no explicit flush"},
        "stages_ms": stages,
        "roofline": roof,
        "roofline_other_kernels": [k for k in others if k],
        "roofline_correlate": {"kernel": "correlate stage (pad + raw SSD + box filters), per direction, against SURVEY 8(d)'s bytes_corr",
                               "bound": "hbm", "achieved": ach_corr, "peak": peak, "unit": "GB/s", "frac": ach_corr / peak if peak else None,
                               "algorithmic_bytes": ab["correlate_per_direction"], "ms": t_corr},
        "gpu_launches": int(launches), "clocks": clocks, "e2e": e2e, "e2e_pipelined": piped,
    }
    if not args.no_cpu_baseline and world == 1:
        import ctypes
        try:    # torch's OpenMP runtime is already up: raise its thread count directly (the oracle shares libgomp)
            ctypes.CDLL("libgomp.so.1").omp_set_num_threads(os.cpu_count())
        except OSError:
            pass
        from oracle import ca_oracle as orc
        orc.build()
        v, dt, desc, cores = cpu_reference_sample(shape, prm, args.niter, 2)
        line["cpu_baseline"] = {"value": v, "unit": "pairs/s", "cores": cores, "kind": "port", "sample": desc, "sample_seconds": dt}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
