#!/usr/bin/env python
"""Benchmark of the raw-trace hot path (BASELINE.json metric, config C2 per GPU).

    python bench.py --gpus N --steps K --warmup W            # ours (one rank per GPU)
    python bench.py --impl reference ...                      # the CPU path on host cores

A step = one pass of the hot path over one synthetic 10-minute Chimera trace per GPU
(2 499 999 600 uint16 samples at 4.17 MHz, 1000 two-level events/s): exact global median
-> fused dequantise + 8-pole 100 kHz Bessel filtfilt -> baseline blocks -> threshold
detection -> CUSUM+ segmentation of every detected event.  `value` is with the raw codes
resident in HBM; `e2e` starts from pinned host memory and ends with the event/level
tables back on the host.  Inputs (5 GB) are far larger than the 126 MB L2, so no explicit
L2 flush is needed between iterations.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

C2_SAMPLES = 2_499_999_600
CUTOFF, ORDER = 100_000.0, 8
THRESHOLD, HYSTERESIS = 5.0, 1.0
BASELINE_BLOCK = 1 << 20
BASELINE_MIN, BASELINE_MAX = 4700.0, 5300.0
EVENT_PAD, MINPOINTS, MAXPOINTS = 100, 8, 100_000
E2E_SHARDS = 16                      # time sub-shards of the streamed end-to-end run (pipeline.StreamingAnalyzer)
CUSUM_DELTA, CUSUM_H = 400.0, 10.0
METRIC = "Msamples/s filtered+CUSUM-segmented"
FILTER_BYTES_PER_SAMPLE = 6.0     # 2 B uint16 read + 4 B float32 written


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def mark(self):
        return time.time()

    def summary(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ts, line in self.rows:
            if ts < t0 - 0.05 or ts > t1 + 0.15:
                continue
            p = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no sample inside the timed region"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------ CPU baseline
def _cpu_chunk(args):
    """The reference's call sequence on one chunk (oracle = port of it; the arithmetic is
    the same scipy/numpy routines the reference calls)."""
    seed, n = args
    from cusumtools_b200 import synth
    from oracle import c_twin, events_oracle as eo, trace_oracle as to
    codes, _ = synth.c1_trace(n=n, n_events=(n - 4000) // synth.EVENT_PERIOD, seed=seed)
    t0 = time.perf_counter()
    data = to.scale_raw_data(codes, synth.CHIMERA_SETTINGS)                  # plot-trace.py:272-287
    y = to.filter_data(data, synth.FS, CUTOFF, ORDER).astype(np.float32)     # plot-trace.py:313-320
    blk = 1 << 16
    c0 = np.float32(0.5 * (BASELINE_MIN + BASELINE_MAX))
    sh = eo.stats_shift(300.0, blk)
    mean, std = eo.baseline_from_stats(*c_twin.block_stats(y, blk, BASELINE_MIN, BASELINE_MAX, c0, sh), c0, sh)
    s, e, _ = c_twin.detect_events(y, blk, *eo.thresholds(mean, std, THRESHOLD, HYSTERESIS))
    w0, w1, typ = eo.event_windows(s, e, n, EVENT_PAD, MINPOINTS, MAXPOINTS)
    ok = typ == 0
    offs = np.concatenate(([0], np.cumsum((w1 - w0)[ok])))
    flat = np.concatenate([y[a:b] for a, b in zip(w0[ok], w1[ok])]) if ok.any() else np.zeros(0, np.float32)
    c_twin.cusum_batch(flat, offs, CUSUM_DELTA, CUSUM_H)
    return time.perf_counter() - t0, n, int(ok.sum())


def cpu_baseline_single(n=1 << 24):
    """1 process / 1 thread, as the reference runs (single-threaded Tk script)."""
    best = None
    for r in range(2):
        dt, nn, ne = _cpu_chunk((r, n))
        best = dt if best is None else min(best, dt)
    return {"value": n / best / 1e6, "unit": "Msamples/s", "cores": 1, "kind": "port",
            "sample": f"{n} samples ({n / 4166666.0:.1f} s of the same synthetic trace), best of 2; "
                      "scale_raw_data + np.pad(median)+filtfilt (scipy) + detection + CUSUM (C twin of the oracle)"}


def run_reference(args):
    """--impl reference: the CPU path on all host cores (chunks of the same workload, one
    process per core, each chunk median-padded on its own — a throughput baseline)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = len(os.sched_getaffinity(0))
    n_chunk = 1 << 23
    times = []
    with mp.get_context("fork").Pool(cores) as pool:
        for it in range(args.warmup + args.steps):
            res = pool.map(_cpu_chunk, [(it * cores + c, n_chunk) for c in range(cores)])
            # all chunks run concurrently; the step takes as long as the slowest one
            # (synthetic-trace generation inside the workers is not part of the path)
            if it >= args.warmup:
                times.append(max(r[0] for r in res))
    total = n_chunk * cores
    ms = 1e3 * float(np.mean(times))
    val = total / (ms / 1e3) / 1e6
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "Msamples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C2: 10-min 4.17 MHz Chimera trace, 8-pole 100 kHz Bessel filtfilt + threshold "
                                   "detection + CUSUM+ (bounded sample per step)", "sample_per_step": total},
            "cpu_baseline": {"value": val, "unit": "Msamples/s", "cores": cores, "kind": "port",
                             "sample": f"{cores} chunks x {n_chunk} samples per step, one process per core; the "
                                       "reference is pure Python over numpy/scipy (nothing to compile into "
                                       "oracle/_ref), so the oracle port of its call sequence is timed"},
            "e2e": {"value": val, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def side_kernels(torch, dev, peak):
    """Kernel-only numbers of the two stages the C2 step does not exercise at their own config shape
    (reported next to the headline, not part of `value`): CUSUM+ on pre-extracted events (C3: 1 M
    events) and the Welch PSD (C4 shape, 2^20-point segments over 2^28 samples); 4 B/sample algorithmic."""
    from cusumtools_b200 import cusum, psd, synth
    out = {}
    x, offs, _ = synth.c3_events_device(1_000_000, dev)
    w0, w1 = offs[:-1].contiguous(), offs[1:].contiguous()
    cusum.cusum_levels(x, w0, w1, delta=CUSUM_DELTA, h=CUSUM_H)
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(3):
        cusum.cusum_levels(x, w0, w1, delta=CUSUM_DELTA, h=CUSUM_H)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 3
    out["cusum_c3_1M_events"] = {"Msamples_per_s": x.numel() / ms / 1e3, "Mevents_per_s": 1_000_000 / ms / 1e3,
                                   "hbm_frac": 4.0 * x.numel() / ms / 1e6 / peak}
    del x, offs, w0, w1
    g = torch.Generator(device=dev); g.manual_seed(3)
    y = torch.randn(1 << 28, generator=g, device=dev) * 24 + 5000
    psd.welch_sums(y, 1 << 20, shift=5000.0)
    torch.cuda.synchronize(); a.record()
    for _ in range(3):
        psd.welch_sums(y, 1 << 20, shift=5000.0)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 3
    out["welch_c4_2p20_segments"] = {"Msamples_per_s": y.numel() / ms / 1e3, "hbm_frac": 4.0 * y.numel() / ms / 1e6 / peak}
    return out


# ------------------------------------------------------------------------------- ours
def run_ours(args):
    import torch
    import torch.distributed as dist
    from cusumtools_b200 import _lib, cusum, detect, filters, pipeline, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = None
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")    # stdout carries exactly one JSON line
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    S = synth.CHIMERA_SETTINGS
    n_own = int(args.samples)
    n_own = n_own // BASELINE_BLOCK * BASELINE_BLOCK if world > 1 else n_own
    halo = pipeline.required_halo(CUTOFF, ORDER, synth.FS, max_event=MAXPOINTS + 2 * EVENT_PAD, block=BASELINE_BLOCK) if world > 1 else 0
    lo_h = halo if rank > 0 else 0
    hi_h = halo if rank < world - 1 else 0
    raw = synth.device_trace(n_own + lo_h + hi_h, dev, seed=1234 + rank, start_index=rank * n_own - lo_h)
    have_cusum = True
    stage_ev = []
    stage_names = ("median", "filter", "baseline", "detect", "cusum")
    an = pipeline.TraceAnalyzer(raw.numel(), S, CUTOFF, ORDER, lo_halo=lo_h, hi_halo=hi_h, threshold=THRESHOLD,
                                hysteresis=HYSTERESIS, baseline_block=BASELINE_BLOCK, baseline_min=BASELINE_MIN,
                                baseline_max=BASELINE_MAX, event_padding=EVENT_PAD, minpoints=MINPOINTS,
                                maxpoints=MAXPOINTS, cusum_delta=CUSUM_DELTA, cusum_h=CUSUM_H, group=group, device=dev)

    def step():
        """One pass of the hot path through the public API (pipeline.TraceAnalyzer.run)."""
        r = an.run(raw)
        return {"starts": r.events.starts, "ends": r.events.ends, "levels": r.levels}

    def staged_step():
        """The same run with a CUDA event after each stage's launches (only for the per-stage
        breakdown; not part of the timed region)."""
        marks = {}

        def hook(name):
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            marks[name] = e

        torch.cuda.synchronize()
        hook("start")
        an.run(raw, stage_hook=hook)
        torch.cuda.synchronize()
        stage_ev.append([marks[k] for k in ("start",) + stage_names])

    def fence():
        torch.cuda.synchronize()
        if group is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, k):
        fence()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a.record()
        r = None
        for _ in range(k):
            r = fn()
        b.record()
        fence()
        wall = (time.perf_counter() - t0) * 1e3
        ms = torch.tensor([a.elapsed_time(b), wall], dtype=torch.float64, device=dev)
        if group is not None:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX, group=group)
        return float(ms[0]), float(ms[1]), r

    sampler = ClockSampler(local) if rank == 0 else None
    for _ in range(args.warmup):
        step()
    _lib.reset_launch_count()
    tm0 = time.time()
    if args.profile_range:
        torch.cuda.profiler.start()
    dev_ms, wall_ms, res = timed(step, args.steps)
    if args.profile_range:
        torch.cuda.profiler.stop()
    tm1 = time.time()
    launches = _lib.launch_count()
    for _ in range(2):
        staged_step()
    clocks = sampler.summary(tm0, tm1) if sampler else None
    n_events = int(res["starts"].numel())
    ev_all = torch.tensor([n_events], dtype=torch.int64, device=dev)
    if group is not None:
        dist.all_reduce(ev_all, group=group)
    stage_ms = {nm: float(np.mean([m[i].elapsed_time(m[i + 1]) for m in stage_ev])) for i, nm in enumerate(stage_names)}
    # the dominant kernel pair alone (forward + backward filter pass, no host round trip in between),
    # timed with CUDA events on the launching stream: the roofline entry
    med = an.last_median if hasattr(an, "last_median") else filters.code_median(raw[lo_h:lo_h + n_own], filters.chimera_bitmask(S))

    def time_pair(stats):
        for _ in range(2):
            filters.dequant_filtfilt(raw, S, CUTOFF, ORDER, median_codes=med, out=an.y, workspace=an.filter_ws, stats=stats)
        fa = torch.cuda.Event(enable_timing=True); fb = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        fa.record()
        for _ in range(3):
            filters.dequant_filtfilt(raw, S, CUTOFF, ORDER, median_codes=med, out=an.y, workspace=an.filter_ws, stats=stats)
        fb.record()
        torch.cuda.synchronize()
        return fa.elapsed_time(fb) / 3

    filt_plain_ms = time_pair(None)
    # the pair as the step runs it: the backward pass also tallies the baseline block sums of its output
    fused_stats = bool(getattr(an, "fuse_stats", False))
    filt_ms = filt_plain_ms
    if fused_stats:
        bl_tmp = detect.new_baseline(raw.numel(), BASELINE_BLOCK, BASELINE_MIN, BASELINE_MAX, dev)
        filt_ms = time_pair(detect.stats_args(bl_tmp, origin=0))
    ms_per_step = dev_ms / args.steps
    total = n_own * world
    value = total / (ms_per_step / 1e3) / 1e6

    # ---- end to end: pinned host -> device, pipeline, tables -> host, every step
    host = torch.empty(raw.numel(), dtype=torch.uint16, pin_memory=True)
    host.copy_(raw)
    torch.cuda.synchronize()
    d2h = [0]

    san = pipeline.StreamingAnalyzer(raw.numel(), S, CUTOFF, ORDER, lo_halo=lo_h, hi_halo=hi_h, shards=E2E_SHARDS,
                                     threshold=THRESHOLD, hysteresis=HYSTERESIS, baseline_block=BASELINE_BLOCK,
                                     baseline_min=BASELINE_MIN, baseline_max=BASELINE_MAX, event_padding=EVENT_PAD,
                                     minpoints=MINPOINTS, maxpoints=MAXPOINTS, cusum_delta=CUSUM_DELTA, cusum_h=CUSUM_H,
                                     group=group, device=dev)
    e2e_info = {}

    def e2e_step():
        # pinned H2D cut into time sub-shards; each is filtered, detected, segmented and its tables are
        # copied back (pinned D2H) while the next pieces arrive; returns after the last synchronisation
        r = san.run_from_host(host)
        d2h[0] = sum(v.nbytes for v in r.tables.values())
        e2e_info.update(events=int(r.tables["starts"].shape[0]), redone=r.redone)
        return r.tables

    e2e_step()
    e_dev_ms, e_wall_ms, _ = timed(e2e_step, max(2, args.steps // 2))
    e_ms = max(e_dev_ms, e_wall_ms) / max(2, args.steps // 2)
    e2e_value = total / (e_ms / 1e3) / 1e6

    if rank != 0:
        if group is not None:
            dist.destroy_process_group()
        return
    peak, peak_kind = peaks()
    ach = FILTER_BYTES_PER_SAMPLE * raw.numel() / (filt_ms / 1e3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get("filter_dram_bytes_per_sample")
            traffic = None if traffic is None else traffic * raw.numel()
    except Exception:
        pass
    line = {
        "metric": METRIC, "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C2: 10-min 4.17 MHz Chimera uint16 trace per GPU, exact median pad + 8-pole 100 kHz "
                               "Bessel filtfilt + baseline blocks + threshold detection"
                               + (" + CUSUM+ on every detected event" if have_cusum else ""),
                   "samples_per_gpu": n_own, "halo_samples": halo, "events_per_step": int(ev_all.item()),
                   "l2": "inputs (5 GB/GPU) larger than L2; no flush needed", "parallelism": f"time-sharded x{world}"},
        "events_per_s": int(ev_all.item()) / (ms_per_step / 1e3),
        "wall_ms_per_step": wall_ms / args.steps,
        "stage_ms": stage_ms,
        "roofline": {"kernel": "ct_filter_fwd_kernel + ct_filter_bwd_kernel (fused dequantise + median pad + zero-phase "
                               "Bessel as two lane-sequential passes" + (", baseline block sums tallied in the backward "
                               "pass's epilogue" if fused_stats else "") + "; timed alone, 3 back-to-back calls)", "bound": "hbm",
                     "achieved": ach, "peak": peak, "peak_kind": peak_kind, "unit": "GB/s", "frac": ach / peak,
                     "traffic": traffic, "kernel_ms": filt_ms,
                     "frac_of_nominal_8000_GBps": ach / 8000.0,      # also against the nominal HBM3e figure
                     "algorithmic_bytes_per_sample": FILTER_BYTES_PER_SAMPLE,
                     "share_of_step": filt_ms / ms_per_step,
                     "filter_only": {"kernel_ms": filt_plain_ms, "frac": FILTER_BYTES_PER_SAMPLE * raw.numel() / (filt_plain_ms / 1e3) / 1e9 / peak,
                                     "note": "the same pair without the fused block sums (the separate ct_block_stats kernel "
                                             "it replaces reads 4 B/sample more and takes 2.2 ms)"},
                     "note": "traffic = ncu dram bytes per launch pair (profiles/): the forward output crosses HBM "
                             "once at half rate (2 B/sample written + 2 B/sample read) on top of the 6 algorithmic bytes"},
        "e2e": {"value": e2e_value, "unit": "Msamples/s", "h2d_bytes_per_step": int(raw.numel() * 2 * world),
                "d2h_bytes_per_step": int(d2h[0] * world), "ms_per_step": e_ms,
                "api": f"pipeline.StreamingAnalyzer.run_from_host ({E2E_SHARDS} time sub-shards overlapped with the pinned H2D copy)",
                "events_rank0": e2e_info.get("events"), "redone": e2e_info.get("redone")},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    if world == 1:
        line["other_kernels"] = side_kernels(torch, dev, peak)
    if world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline_single()
    emit(line)
    if group is not None:
        dist.destroy_process_group()


_RESULT_FD = None


def emit(line: dict) -> None:
    """The one JSON line, written to the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def main():
    # stdout must carry exactly one JSON line: libraries that write to file descriptor 1 (NCCL prints its
    # version banner there) are sent to stderr for the whole run, the result goes to the saved descriptor
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--samples", type=int, default=C2_SAMPLES, help="samples per GPU (default: config C2)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--profile-range", action="store_true",
                    help="bracket the timed region with cudaProfilerStart/Stop (ncu --profile-from-start off)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
