"""The raw-trace hot path end to end on one GPU or time-sharded over several.

    raw Chimera codes -> [exact global median] -> fused dequantise + Bessel filtfilt
                      -> baseline blocks -> threshold/hysteresis events

One process per GPU.  A rank owns the samples [lo, hi) of the global trace and is handed
them together with `halo` extra samples on each side (read from the shared file / host
buffer, not exchanged: SURVEY.md section 8e).  The only collectives are
  * all_reduce(SUM) of the sampled code histogram and of the 9 window counters that give
    the GLOBAL median pad value (so every rank subtracts the same constant), and
  * all_gather of per-rank event counts (global event ids); event rows stay rank-local.
With `group=None` nothing is communicated and the functions are the single-GPU path.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib, cusum, detect, filters
from .design import bessel_lowpass


@dataclass
class TraceResult:
    filtered: torch.Tensor          # float32 device, the rank's owned samples only
    baseline: detect.Baseline
    events: detect.EventList        # indices relative to the rank's first owned sample
    pad_value: float
    median_codes: tuple[int, int]
    first_event_id: int = 0         # global id of this rank's first event
    total_events: int = 0


def _world(group) -> int:
    import torch.distributed as dist
    return dist.get_world_size(group)


def _all_reduce_(t: torch.Tensor, group) -> torch.Tensor:
    if group is not None:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


@dataclass
class MedianPlan:
    """State of the exact-median search between its two phases (estimate, verify)."""
    n: int                      # samples of all ranks
    k1: int                     # ranks of the two middle order statistics
    k2: int
    step: int                   # spacing of the masked codes
    shift: int
    est: int                    # estimated median code (a multiple of step)
    lo: int                     # first code of the 8-code verification window
    exact: tuple[int, int] | None = None     # already exact (small traces: full histogram)
    se: float = float("inf")    # standard error of the estimate in code steps: sqrt(sampled) / (2 x samples at the estimate)


def median_estimate(n_local: int, mask: int, hist_fn, group=None, device=None, n_sampled: int | None = None) -> MedianPlan:
    """Phase 1: a strided-sample histogram (summed over `group`) locates the median code.  `n_sampled`:
    the histogram covers only that many of the rank's samples (streaming: the first chunk), so it
    is an estimate even for small traces."""
    shift = 0
    while shift < 16 and not (mask >> shift) & 1:
        shift += 1
    step = 1 << shift
    if group is None:
        n = int(n_local)
        if n == 0:
            raise ValueError("median of an empty trace")
        stride = max(1, (n if n_sampled is None else int(n_sampled)) // (1 << 20))  # ~1 M samples: median s.e. < 0.1 code
        hist = hist_fn(stride)
    else:
        # every rank samples with the stride of ITS share (ranks own equal shares), and the sample count rides in the
        # same all_reduce as the histogram: one collective for the phase
        stride = max(1, (int(n_local) if n_sampled is None else int(n_sampled)) * _world(group) // (1 << 20))
        h = hist_fn(stride).to(torch.int64)
        packed = _all_reduce_(torch.cat((h, torch.tensor([int(n_local)], dtype=torch.int64, device=h.device))), group)
        hist, n = packed[:-1], None                        # (the total comes back with the rank search: one read)
    exact = stride == 1 and n_sampled is None
    if group is not None and (exact or not hist.is_cuda):
        n = int(packed[-1].item())
    if exact:
        if n == 0:
            raise ValueError("median of an empty trace")
        k1, k2 = (n - 1) // 2, n // 2
        cdf = torch.cumsum(hist.to(torch.int64), 0)
        want = torch.tensor([k1 + 1, k2 + 1], dtype=torch.int64, device=cdf.device)
        c1, c2 = (int(v) for v in torch.searchsorted(cdf, want).tolist())
        return MedianPlan(n, k1, k2, step, shift, c1, max(0, c1 - 3 * step), exact=(c1, c2))
    # the rank search runs where the histogram lives: 24 (32) bytes come back instead of 65 536 bins
    if hist.is_cuda and hist.dtype in (torch.int32, torch.int64) and hist.is_contiguous():
        out = torch.empty(3, dtype=torch.int64, device=hist.device)
        with torch.cuda.device(hist.device):
            rc = _lib.lib().ct_hist_rank(hist.data_ptr(), int(hist.dtype == torch.int64), out.data_ptr(),
                                         C.c_void_p(torch.cuda.current_stream(hist.device).cuda_stream))
        _lib.check(rc, "ct_hist_rank")
        vals = (torch.cat((out, packed[-1:])) if n is None else out).tolist()
        i, dens, total = (int(v) for v in vals[:3])
        if n is None:
            n = int(vals[3])
    else:                                                  # (CPU tensors: the gloo tests of the host logic)
        hist = hist.to(torch.int64)
        cdf = torch.cumsum(hist, 0)
        idx = torch.searchsorted(cdf, (cdf[-1:] + 1) // 2).clamp_(max=hist.numel() - 1)
        i, dens, total = (int(v) for v in torch.cat((idx, hist[idx], cdf[-1:])).tolist())
    if n == 0:
        raise ValueError("median of an empty trace")
    k1, k2 = (n - 1) // 2, n // 2
    est = (i >> shift) * step
    # the sample median of m samples with density f at the median has s.e. 1 / (2 f sqrt(m)); f = dens / m per code step
    se = 0.5 * float(np.sqrt(max(total, 1))) / max(dens, 1)
    return MedianPlan(n, k1, k2, step, shift, est, max(0, est - 3 * step), se=se)


def median_verify(plan: MedianPlan, counts9, group=None):
    """Phase 2: the exact counts of the codes below / inside the window (summed over `group`) either pin
    the two middle order statistics or say which way the window has to move (returns None, new lo)."""
    c = _all_reduce_(counts9.to(torch.int64), group).cpu().numpy().astype(np.int64)
    below, cw = int(c[0]), int(c[0]) + np.cumsum(c[1:])
    if below <= plan.k1 and plan.k2 < cw[-1]:
        return (plan.lo + int(np.searchsorted(cw, plan.k1 + 1)) * plan.step,
                plan.lo + int(np.searchsorted(cw, plan.k2 + 1)) * plan.step), plan.lo
    return None, (max(0, plan.lo - 6 * plan.step) if plan.k1 < below else plan.lo + 6 * plan.step)


def median_search(n_local: int, mask: int, hist_fn, count_fn, group=None, device=None, plan: MedianPlan | None = None,
                  first_counts=None) -> tuple[int, int]:
    """Host side of the exact global median: the two middle order statistics of the masked
    codes of ALL ranks.  `hist_fn(stride)` returns this rank's strided-sample histogram
    (int tensor[65536]) and `count_fn(lo, step)` its 9 window counters (#codes < lo, then
    #codes == lo + i*step); both are summed over `group` here, so every rank takes the same
    decisions and returns the same pair.  (The device kernels are passed in, which is what
    lets the world-size-2 gloo test drive this logic on CPU.)  `plan` / `first_counts`: an
    estimate already made and the window counts a fused kernel already produced for it."""
    if plan is None:
        plan = median_estimate(n_local, mask, hist_fn, group, device)
    if plan.exact is not None:
        return plan.exact
    counts = first_counts
    for attempt in range(16):
        if counts is None:
            counts = count_fn(plan.lo, plan.step)
        pair, plan.lo = median_verify(plan, counts, group)
        if pair is not None:
            return pair
        if attempt == 0:
            # the window missed: the estimate came from too small or too local a sample (streaming: the first
            # piece of a drifting trace).  Re-estimate from a strided histogram of ALL the data before stepping.
            again = median_estimate(n_local, mask, hist_fn, group, device)
            if again.exact is not None:
                return again.exact
            plan.lo = again.lo                 # (plan.est stays: the filter has already subtracted it)
        counts = None
    cdf = np.cumsum(_all_reduce_(hist_fn(1).to(torch.int64), group).cpu().numpy())   # pathological distribution
    return int(np.searchsorted(cdf, plan.k1 + 1)), int(np.searchsorted(cdf, plan.k2 + 1))


def _median_kernels(raw_owned: torch.Tensor, mask: int):
    """(hist_fn, count_fn) over this rank's owned codes."""
    L = _lib.lib()
    dev = raw_owned.device
    n_local = raw_owned.numel()
    st = filters._stream_ptr(raw_owned)

    def hist_fn(stride):
        h = torch.zeros(65536, dtype=torch.int32, device=dev)
        if n_local:
            _lib.check(L.ct_hist_sampled_u16(raw_owned.data_ptr(), n_local, stride, mask, h.data_ptr(), st), "ct_hist_sampled_u16")
        return h

    def count_fn(lo, step):
        cnt = torch.zeros(9, dtype=torch.int64, device=dev)
        if n_local:
            _lib.check(L.ct_count_window_u16(raw_owned.data_ptr(), n_local, mask, lo, step, cnt.data_ptr(), st), "ct_count_window_u16")
        return cnt

    return hist_fn, count_fn


def global_code_median(raw_owned: torch.Tensor, mask: int, group=None) -> tuple[int, int]:
    """Exact median (two middle order statistics) of the masked codes of the WHOLE trace
    when each rank passes its owned samples; identical to filters.code_median for one rank."""
    if group is None:
        return filters.code_median(raw_owned, mask)
    hist_fn, count_fn = _median_kernels(raw_owned, mask)
    return median_search(raw_owned.numel(), mask, hist_fn, count_fn, group, raw_owned.device)


def event_id_offsets(n_local_events: int, group=None, device=None) -> tuple[int, int]:
    """(global id of this rank's first event, total events): ranks own consecutive time
    shards, so ids follow rank order."""
    if group is None:
        return 0, int(n_local_events)
    import torch.distributed as dist
    ws, rk = dist.get_world_size(group), dist.get_rank(group)
    counts = torch.zeros(ws, dtype=torch.int64, device=device)
    counts[rk] = int(n_local_events)
    c = _all_reduce_(counts, group).cpu().numpy()
    return int(c[:rk].sum()), int(c.sum())


def event_ids_and_status(n_local_events: int, status: int, group=None, device=None) -> tuple[int, int, int]:
    """`event_id_offsets` with a status flag riding in the same all_reduce: (first id, total events, sum of the
    ranks' flags).  A rank that hit an error passes status != 0 INSTEAD of raising before the collective, so that
    every rank takes the same decision afterwards (raise together, fall back together) and nobody is left waiting
    in a collective its peers never enter."""
    if group is None:
        return 0, int(n_local_events), int(status)
    import torch.distributed as dist
    ws, rk = dist.get_world_size(group), dist.get_rank(group)
    v = torch.zeros(ws + 1, dtype=torch.int64, device=device)
    v[rk] = int(n_local_events)
    v[ws] = int(status != 0)
    c = _all_reduce_(v, group).cpu().numpy()
    return int(c[:rk].sum()), int(c[:ws].sum()), int(c[ws])


def required_halo(cutoff: float, order: int, samplerate: float, max_event: int, padding: int = 1000,
                  eps: float = filters.DEFAULT_HALO_EPS, block: int = 1) -> int:
    """Samples a rank must read beyond each end of its owned range: IIR warm-up on both
    sides (forward and backward pass) plus one maximal event so that an event straddling
    a shard boundary is seen whole by the rank that owns its start, rounded up to a multiple
    of the baseline `block` (the left halo keeps every rank on the global block grid)."""
    d = bessel_lowpass(int(order), 2.0 * float(cutoff) / float(samplerate))
    h = 2 * max(1, d.impulse_tail(eps)) + int(max_event)
    return -(-h // int(block)) * int(block)


@dataclass
class AnalysisResult:
    """Everything one pass of the hot path over a (shard of a) trace produces; all tensors
    are device-resident views into the analyzer's buffers (valid until its next run)."""
    filtered: torch.Tensor          # float32, the rank's owned samples
    detect_trace: torch.Tensor      # float32, [left halo | owned | right halo]: what the windows refer to
    lo_halo: int                    # samples of detect_trace before the first owned one
    baseline: detect.Baseline
    events: detect.EventList        # indices relative to the rank's first owned sample
    win_start: torch.Tensor         # int64 [E]  CUSUM+ windows [start - padding, end + padding) in detect_trace
    win_end: torch.Tensor           # int64 [E]
    types: torch.Tensor             # int32 [E]  rate.csv type code (0 accepted, >1 rejected)
    levels: "cusum.LevelTable | None"
    pad_value: float
    median_codes: tuple[int, int]
    first_event_id: int = 0
    total_events: int = 0
    intra: "tuple[torch.Tensor, torch.Tensor] | None" = None     # (count int32 [E], pairs int32 [E, 2K]) crossings


class TraceAnalyzer:
    """Stages 1-3 over a time shard with every buffer allocated once and a single host
    synchronisation at the end of the step (the median needs two small ones up front).

    `raw_ext` = [lo_halo | owned | hi_halo] codes.  The filter runs over the extended range
    (its constant pad only matters at the true ends of the trace), and so do the baseline
    blocks and the detector: the state at the first owned sample is then the true one (an
    event in progress at a shard boundary is not seen twice) and an event that starts in the
    owned range and ends in the right halo is completed.  Events are owned by the rank whose
    owned range holds their start.  `lo_halo` must be a multiple of the baseline block so
    that the block grid of every rank is the global one (`required_halo(..., block=)`)."""

    def __init__(self, n_ext: int, settings, cutoff: float, order: int = 8, *, lo_halo: int = 0, hi_halo: int = 0,
                 threshold: float = 5.0, hysteresis: float = 1.0, baseline_block: int = detect.DEFAULT_BASELINE_BLOCK,
                 baseline_min: float, baseline_max: float, padding: int = 1000, event_padding: int = 100,
                 minpoints: int = 8, maxpoints: int = 100_000, cusum_delta: float | None = None,
                 cusum_h: float | None = None, max_levels: int = cusum.DEFAULT_MAX_LEVELS,
                 event_capacity: int | None = None, group=None, device="cuda", fuse_stats: bool = True,
                 fused_count: bool = True, intra_threshold: float = 0.0, intra_hysteresis: float = 0.0,
                 max_crossings: int = 8, halos_clipped_by_trace_ends: bool = False):
        self.n_ext, self.lo_halo, self.hi_halo = int(n_ext), int(lo_halo), int(hi_halo)
        self.n_own = self.n_ext - self.lo_halo - self.hi_halo
        self.n_det = self.n_ext
        if self.lo_halo % int(baseline_block):
            raise ValueError("lo_halo must be a multiple of the baseline block (pipeline.required_halo(..., block=))")
        # a halo (0 = a true end of the trace) must cover the IIR warm-up of both passes and one maximal event window
        need = required_halo(cutoff, order, float(np.floor(np.squeeze(settings["ADCSAMPLERATE"]))),
                             int(maxpoints) + 2 * int(event_padding), int(padding))
        # (`halos_clipped_by_trace_ends`: the caller cut full halos at the true ends of the trace, where nothing is missing)
        for name, h in (("lo_halo", self.lo_halo), ("hi_halo", self.hi_halo)):
            if 0 < h < need and not halos_clipped_by_trace_ends:
                raise ValueError(f"{name} = {h} samples is shorter than the {need} this configuration needs "
                                 "(pipeline.required_halo(cutoff, order, fs, maxpoints + 2 * event_padding))")
        self.settings, self.cutoff, self.order, self.padding = settings, float(cutoff), int(order), int(padding)
        self.threshold, self.hysteresis = float(threshold), float(hysteresis)
        self.block, self.bmin, self.bmax = int(baseline_block), float(baseline_min), float(baseline_max)
        self.event_padding, self.minpoints, self.maxpoints = int(event_padding), int(minpoints), int(maxpoints)
        self.delta, self.h, self.max_levels = cusum_delta, cusum_h, int(max_levels)
        # intra-event threshold crossings (readevents.py:1340-1343, 1363-1367); 0 = off, as in summary.txt
        self.intra_threshold, self.intra_hysteresis = float(intra_threshold), float(intra_hysteresis)
        self.max_crossings = int(max_crossings)
        self.group = group
        self.device = torch.device(device)
        self.mask = filters.chimera_bitmask(settings)
        L = _lib.lib()
        if self.block % L.ct_detect_run():
            raise ValueError(f"baseline block must be a multiple of {L.ct_detect_run()} samples")
        self.y = torch.empty(self.n_ext, dtype=torch.float32, device=self.device)
        self.design = bessel_lowpass(self.order, 2.0 * self.cutoff / float(np.floor(np.squeeze(settings["ADCSAMPLERATE"]))))
        # baseline block sums ride on the filter's epilogue when the block is a whole number of its warp groups
        self.fuse_stats = (bool(fuse_stats) and self.block >= 65536
                           and self.block % filters.stats_granule(self.n_ext, self.padding, self.design) == 0)
        self.fused_count = bool(fused_count)     # the exact-median window tally rides on the forward pass (+0.35 ms against 0.79 ms as its own kernel on C2)
        self.filter_ws = None
        self.minmax = None                         # (min, max) of every 64-sample chunk of y, left by the backward pass
        self.H = max(1, self.design.impulse_tail(filters.DEFAULT_HALO_EPS))
        self.ws_bytes = int(L.ct_detect_workspace_bytes(self.n_det))
        self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=self.device)
        self._alloc_events(int(event_capacity) if event_capacity else max(1024, self.n_det // 2048))

    def _alloc_events(self, cap: int) -> None:
        dev, ML = self.device, self.max_levels
        self.cap = cap
        self.starts = torch.empty(cap, dtype=torch.int64, device=dev)
        self.ends = torch.empty(cap, dtype=torch.int64, device=dev)
        self.w0 = torch.empty(cap, dtype=torch.int64, device=dev)
        self.w1 = torch.empty(cap, dtype=torch.int64, device=dev)
        self.typ = torch.empty(cap, dtype=torch.int32, device=dev)
        self.scalars = torch.zeros(4, dtype=torch.int64, device=dev)     # n_starts, n_ends, n_kept, first kept index
        self.median_result = torch.zeros(4, dtype=torch.int32, device=dev)   # CtMedianResult: code1, code2, pad_x (float bits), status
        if self.intra_threshold > 0:
            self.ic = torch.zeros(cap, dtype=torch.int32, device=dev)
            self.ip = torch.full((cap, 2 * self.max_crossings), -1, dtype=torch.int32, device=dev)
        if self.delta is not None:
            self.nl = torch.empty(cap, dtype=torch.int32, device=dev)
            self.ed = torch.empty((cap, ML + 1), dtype=torch.int32, device=dev)
            self.mu = torch.empty((cap, ML), dtype=torch.float64, device=dev)
            self.sd = torch.empty((cap, ML), dtype=torch.float64, device=dev)
            self.ov = torch.empty(cap, dtype=torch.uint8, device=dev)
            self.cws_bytes = int(_lib.lib().ct_cusum_workspace_bytes(cap))
            self.cws = torch.empty((self.cws_bytes + 7) // 8, dtype=torch.int64, device=dev)

    def run_from_host(self, host_codes: torch.Tensor, chunks: int = 16, stage_hook=None) -> AnalysisResult:
        """`run` with the codes still in (pinned) host memory: the host->device copy is cut into
        `chunks` pieces on a copy stream and the forward filter pass (with the median count on its
        side) follows the copy chunk by chunk, so only the work after the forward pass is left when the
        last byte arrives.  Possible because the forward pass needs the median only approximately:
        the estimate comes from the first chunk, the exact value from the fused count."""
        if host_codes.numel() != self.n_ext or host_codes.dtype not in (torch.uint16, torch.int16) or host_codes.is_cuda:
            raise ValueError("host_codes must be a CPU uint16/int16 tensor of the planned length")
        if getattr(self, "raw_dev", None) is None:
            self.raw_dev = torch.empty(self.n_ext, dtype=host_codes.dtype, device=self.device)
            self.copy_stream = torch.cuda.Stream(device=self.device)
        gran = 1 << 20
        per = max(gran, -(-self.n_ext // int(chunks) // gran) * gran)
        bounds = list(range(0, self.n_ext, per)) + [self.n_ext]
        events = []
        cur = torch.cuda.current_stream(self.device)
        self.copy_stream.wait_stream(cur)                      # the previous step may still read raw_dev
        with torch.cuda.stream(self.copy_stream):
            for a, b in zip(bounds[:-1], bounds[1:]):
                self.raw_dev[a:b].copy_(host_codes[a:b], non_blocking=True)
                e = torch.cuda.Event()
                e.record(self.copy_stream)
                events.append(e)
        return self.run(self.raw_dev, stage_hook=stage_hook, _arrivals=(bounds, events))

    def run(self, raw_ext: torch.Tensor, stage_hook=None, _arrivals=None, _fixed=None) -> AnalysisResult:
        # the library keys its per-device state on the CURRENT device: make it the tensors' device
        with torch.cuda.device(self.device):
            return self._run(raw_ext, stage_hook, _arrivals, _fixed)

    def _run(self, raw_ext: torch.Tensor, stage_hook=None, _arrivals=None, _fixed=None, _host_median: bool = False) -> AnalysisResult:
        """One pass of stages 1-3.  `stage_hook(name)` (optional) is called after the launches of each
        stage have been enqueued: 'median', 'filter', 'baseline', 'detect', 'cusum' (profiling only).
        `_fixed = (MedianPlan, pad_x, (c1, c2))` (StreamingAnalyzer) skips the median stages: the filter
        subtracts `plan.est` and pads with `pad_x` = median - estimate."""
        hook = stage_hook or (lambda name: None)
        if raw_ext.numel() != self.n_ext:
            raise ValueError(f"analyzer was planned for {self.n_ext} samples, got {raw_ext.numel()}")
        L = _lib.lib()
        lo, n_own = self.lo_halo, self.n_own
        owned = raw_ext[lo:lo + n_own]
        st = filters._stream_ptr(raw_ext)
        cur = torch.cuda.current_stream(self.device)
        # ---- median, phase 1: estimate (subtraction constant of the filter, verification window)
        hist_fn, count_fn = _median_kernels(owned, self.mask)
        if _fixed is not None:
            plan = _fixed[0]
        elif _arrivals is None:
            plan = median_estimate(n_own, self.mask, hist_fn, self.group, self.device)
        else:                                                  # only the first chunk has arrived: estimate from it
            bounds, events = _arrivals
            cur.wait_event(events[0])
            first = raw_ext[lo:max(lo + 1, min(bounds[1], lo + n_own))]
            plan = median_estimate(n_own, self.mask, _median_kernels(first, self.mask)[0], self.group, self.device,
                                   n_sampled=first.numel())
        hook("median")
        if self.filter_ws is None:
            need = int(L.ct_filtfilt_workspace_bytes(self.n_ext, self.padding, self.H))
            self.filter_ws = torch.empty(need, dtype=torch.uint8, device=self.device)
            self.minmax = torch.empty(2 * int(L.ct_filter_summary_count(self.n_ext, self.padding, self.H)), dtype=torch.float32,
                                      device=self.device)
        coef = filters.make_coef(self.design)
        alpha, _ = filters.chimera_affine(self.settings)
        origin = 0
        # ---- forward pass with the estimate; it tallies the window counts of the owned codes on the side
        counts = torch.zeros(9, dtype=torch.int64, device=self.device) if _fixed is None else None
        fused = plan.exact is None and _fixed is None
        pieces = [(0, 0, self.n_ext)] if _arrivals is None else [(2, a, b) for a, b in zip(_arrivals[0][:-1], _arrivals[0][1:])]
        pad_first = float(_fixed[1]) if _fixed is not None else 0.0
        # The estimate of a resident trace comes from a sample of ALL of it, on every rank the same; its standard error
        # follows from the density at the estimate (MedianPlan.se, in code steps).  Up to 0.8 the eight-code window
        # est - 3 .. est + 4 holds the median (4 sigma) and is verified on the device; the tally rides on the forward pass
        # (4.1 instructions per code) or runs as its own kernel (four codes est - 1 .. est + 2 when se <= 0.25).  Beyond
        # that - a drifting baseline - or from a partial trace (streaming): the host loop.
        device_ok = fused and _arrivals is None and plan.se <= 0.8
        ride = device_ok and self.fused_count and plan.step <= 4
        narrow = device_ok and not ride and plan.se <= 0.25
        nbins = 4 if narrow else 8
        if narrow:
            plan.lo = max(0, plan.est - plan.step)
        for i, (part, a, b) in enumerate(pieces):
            if _arrivals is not None:
                cur.wait_event(_arrivals[1][i])
            rc = L.ct_filter_forward_u16(raw_ext.data_ptr(), self.n_ext, self.padding, float(plan.est), self.mask, pad_first,
                                         C.byref(coef), self.H, origin, part, plan.lo, plan.step, lo, lo + n_own,
                                         counts.data_ptr() if ride else None, a, b,
                                         self.filter_ws.data_ptr(), self.filter_ws.numel(), st)
            _lib.check(rc, "ct_filter_forward_u16")
            if fused and not ride:                 # the window count of this piece's owned codes as its own kernel
                ca, cb = max(a, lo), min(b, lo + n_own)
                if cb > ca:
                    count = L.ct_count_window4_u16 if narrow else L.ct_count_window_u16
                    rc = count(raw_ext[ca:cb].data_ptr(), cb - ca, self.mask, plan.lo, plan.step, counts.data_ptr(), st)
                    _lib.check(rc, "ct_count_window_u16")
        # everything the backward launch needs that does not depend on the exact median is prepared BEFORE the host waits
        # for the counts: the GPU idles from the end of the forward pass to the backward launch
        bl = detect.new_baseline(self.n_det, self.block, self.bmin, self.bmax, self.device) if self.fuse_stats else None
        stats = detect.stats_args(bl, origin=0) if bl is not None else None
        offset = float(filters.scale_codes_host(np.array([plan.est], dtype=np.uint16), self.settings)[0])
        # ---- median, phase 2: exact order statistics
        on_device = device_ok and not _host_median
        self.last_median_route = "device" if on_device else ("host after a device miss" if _host_median else "host")
        if on_device:
            # no host round trip between the passes: one thread turns the (summed) counts into the two middle codes and the
            # pad of the filter ends, the end groups re-run with it (their CTAs return at once when the estimate was the
            # median), and the backward pass follows; the codes reach the host with the step's one synchronisation.  If
            # the window missed them (status != 0) the step is redone the host-driven way.
            _all_reduce_(counts, self.group)
            res = self.median_result
            rc = L.ct_median_verify(counts.data_ptr(), nbins, plan.k1, plan.k2, plan.lo, plan.step, float(plan.est),
                                    res.data_ptr(), st)
            _lib.check(rc, "ct_median_verify")
            rc = L.ct_filter_forward_ends_u16(raw_ext.data_ptr(), self.n_ext, self.padding, float(plan.est), self.mask,
                                              res[2:].data_ptr(), C.byref(coef), self.H, origin,
                                              self.filter_ws.data_ptr(), self.filter_ws.numel(), st)
            _lib.check(rc, "ct_filter_forward_ends_u16")
            c1 = c2 = None
        else:
            if _fixed is not None:
                c1, c2 = _fixed[2]
            else:
                c1, c2 = median_search(n_own, self.mask, hist_fn, count_fn, self.group, self.device, plan=plan,
                                       first_counts=counts if fused else None)
            pad_x = 0.5 * (c1 + c2) - plan.est if _fixed is None else 0.0
            if pad_x != 0.0:        # the pad holds median - estimate: redo the groups at the two ends of the trace
                rc = L.ct_filter_forward_u16(raw_ext.data_ptr(), self.n_ext, self.padding, float(plan.est), self.mask,
                                             float(pad_x), C.byref(coef), self.H, origin, 1, 0, 1, 0, 0, None, 0, 0,
                                             self.filter_ws.data_ptr(), self.filter_ws.numel(), st)
                _lib.check(rc, "ct_filter_forward_u16")
        y = self.y
        rc = L.ct_filter_backward(self.n_ext, self.padding, float(plan.est), float(alpha), offset, C.byref(coef), self.H, origin,
                                  y.data_ptr(), self.filter_ws.data_ptr(), self.filter_ws.numel(),
                                  C.byref(stats) if stats is not None else None, self.minmax.data_ptr(), st)
        _lib.check(rc, "ct_filter_backward")
        yd = y
        st = filters._stream_ptr(y)
        hook("filter")
        if bl is not None:
            detect.finish_baseline(bl, self.threshold, self.hysteresis)
        else:
            bl = detect.baseline_blocks(yd, self.block, self.bmin, self.bmax, threshold=self.threshold,
                                        hysteresis=self.hysteresis)
        sign, ts, te = bl.device_lines(self.device)
        hook("baseline")
        while True:
            sc = self.scalars
            rc = L.ct_detect_f32(yd.data_ptr(), self.n_det, self.block, sign.data_ptr(), ts.data_ptr(), te.data_ptr(), 0,
                                 self.minmax.data_ptr(), 0, self.ws.data_ptr(), self.ws_bytes, self.starts.data_ptr(),
                                 self.ends.data_ptr(), self.cap, sc[0:].data_ptr(), st)
            _lib.check(rc, "ct_detect_f32")
            rc = L.ct_event_windows(self.starts.data_ptr(), self.ends.data_ptr(), sc[0:].data_ptr(), self.cap, self.n_det,
                                    lo, lo + n_own, self.event_padding, self.minpoints, self.maxpoints, self.w0.data_ptr(),
                                    self.w1.data_ptr(), self.typ.data_ptr(), sc[2:].data_ptr(), st)
            _lib.check(rc, "ct_event_windows")
            hook("detect")
            if self.delta is not None:
                rc = L.ct_cusum_batch_dev(yd.data_ptr(), self.n_det, self.w0.data_ptr(), self.w1.data_ptr(),
                                          self.typ.data_ptr(), sc[2:].data_ptr(), self.cap, float(self.delta),
                                          float(self.h), self.max_levels, self.nl.data_ptr(), self.ed.data_ptr(),
                                          self.mu.data_ptr(), self.sd.data_ptr(), self.ov.data_ptr(), self.cws.data_ptr(),
                                          self.cws_bytes, st)
                _lib.check(rc, "ct_cusum_batch_dev")
            if self.intra_threshold > 0:
                # the block of the event start: win_start + padding (windows are compacted, the start list is not)
                detect.intra_crossings(yd, self.w0, self.w1, self.w0 + self.event_padding, bl, self.intra_threshold,
                                       self.intra_hysteresis, max_pairs=self.max_crossings, n_events_dev=sc[2:],
                                       out=(self.ic, self.ip))
            hook("cusum")
            tail = (self.median_result.to(torch.int64),) if on_device else ()
            host = torch.cat((sc[:4], bl.dev["status"].to(torch.int64)) + tail).cpu().numpy()   # the step's one sync
            if on_device:
                if int(host[8]) != 0:                # the four-code window missed the median (every rank sees the same counts)
                    return self._run(raw_ext, stage_hook, _arrivals, _fixed, _host_median=True)
                c1, c2 = int(host[5]), int(host[6])
            ns, ne, nk, i0 = int(host[0]), int(host[1]), int(host[2]), int(host[3])
            if max(ns, ne) <= self.cap:
                break
            self._alloc_events(max(ns, ne))          # more events than planned: grow and redo stages 2-3
        self.last_median = (c1, c2)
        pad_value = float(np.median(filters.scale_codes_host(np.array([c1, c2], dtype=np.uint16), self.settings)))
        status = int(host[4]) != 0
        if _fixed is None:
            # the error decision is collective: the flag rides with the event counts, then every rank raises
            first_id, total, bad = event_ids_and_status(0 if status else nk, int(status), self.group, self.device)
        else:
            first_id, total, bad = 0, nk, int(status)
        if bad:
            raise ValueError("no baseline block has enough samples inside [baseline_min, baseline_max]"
                             + ("" if status else " (on another rank)"))
        bl._checked = True
        open_start = -1
        if ns > ne:                                   # an event still open at the end of the data
            o = int(self.starts[ne].item()) - lo
            open_start = o if 0 <= o < n_own else -1
        # event indices are reported relative to the first owned sample
        ev = detect.EventList(self.starts[i0:i0 + nk] - lo, self.ends[i0:i0 + nk] - lo, open_start)
        lv = None
        if self.delta is not None:
            lv = cusum.LevelTable(self.nl[:nk], self.ed[:nk], self.mu[:nk], self.sd[:nk], self.ov[:nk], self.max_levels)
        return AnalysisResult(filtered=y[lo:lo + n_own], detect_trace=yd, lo_halo=lo, baseline=bl, events=ev,
                              win_start=self.w0[:nk], win_end=self.w1[:nk], types=self.typ[:nk], levels=lv,
                              pad_value=pad_value, median_codes=(c1, c2), first_event_id=first_id, total_events=total,
                              intra=(self.ic[:nk], self.ip[:nk]) if self.intra_threshold > 0 else None)


    def tables_to_host(self, r: AnalysisResult) -> dict:
        """Event and level tables of `r` as numpy arrays: device -> pinned host buffers
        (allocated once at capacity) with one synchronisation; the arrays alias the pinned
        buffers and are valid until the next call."""
        nk = int(r.events.starts.numel())
        if getattr(self, "_pinned_cap", 0) < self.cap:
            self._pinned = {}
            self._pinned_cap = self.cap
        src = {"starts": r.events.starts, "ends": r.events.ends, "types": r.types}
        if r.levels is not None:
            src.update(n_levels=r.levels.n_levels, edges=r.levels.edges, mean=r.levels.mean, std=r.levels.std,
                       overflow=r.levels.overflow)
        if r.intra is not None:
            src.update(intra_count=r.intra[0], intra_pairs=r.intra[1])
        out = {}
        for k, t in src.items():
            if k not in self._pinned:
                self._pinned[k] = torch.empty((self.cap,) + tuple(t.shape[1:]), dtype=t.dtype, pin_memory=True)
            h = self._pinned[k][:nk]
            h.copy_(t, non_blocking=True)
            out[k] = h
        torch.cuda.current_stream(self.device).synchronize()
        return {k: v.numpy() for k, v in out.items()}


@dataclass
class StreamResult:
    """What `StreamingAnalyzer.run_from_host` leaves behind: the filtered trace and the baseline
    table on the device, the event / level tables already in (pinned) host memory."""
    filtered: torch.Tensor          # float32 CUDA, the rank's owned samples
    baseline: detect.Baseline       # blocks of the owned range
    tables: dict                    # numpy: starts, ends (relative to the first owned sample), types[, n_levels, edges, mean, std, overflow]
    pad_value: float
    median_codes: tuple[int, int]
    first_event_id: int = 0
    total_events: int = 0
    redone: str = ""                # "", "first" (first sub-shard redone with the exact pad) or "whole" (fallback)


class _HostFeed:
    """Pieces of a trace that already sits in (pinned) host memory: every copy is queued up front."""

    def __init__(self, host_codes: torch.Tensor):
        self.host = host_codes

    def start(self, an: "StreamingAnalyzer", pieces, cur) -> None:
        self.events = []
        an.copy_stream.wait_stream(cur)                        # the previous step may still read raw_dev
        with torch.cuda.stream(an.copy_stream):
            for a, b in pieces:
                if b > a:
                    an.raw_dev[a:b].copy_(self.host[a:b], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(an.copy_stream)
                self.events.append(ev)

    def wait(self, i: int, cur) -> None:
        cur.wait_event(self.events[i])

    def finish(self) -> None:
        pass

    def whole(self, an: "StreamingAnalyzer") -> torch.Tensor:
        return self.host


class _FileFeed:
    """Pieces read from the `.log` series on a worker thread (loader.WindowReader: several reader threads fill
    disjoint parts of a pinned slab with preadv, GIL released), each handed to the copy engine as soon as it is in
    memory; `slabs` pinned slabs rotate, so reading piece i+1 overlaps the host->device copy of piece i and the
    kernels of piece i-1 (plot-trace.py:230-299 reads the whole window before anything else happens)."""

    def __init__(self, reader, threads: int = 8, slabs: int = 3):
        self.reader, self.threads, self.nslabs = reader, int(threads), int(slabs)
        self.error = None

    def start(self, an: "StreamingAnalyzer", pieces, cur) -> None:
        import threading
        need = max((b - a for a, b in pieces), default=1)
        pool = getattr(an, "_slabs", None)
        if pool is None or len(pool) != self.nslabs or pool[0].numel() < need:
            pool = an._slabs = [torch.empty(need, dtype=an.raw_dev.dtype, pin_memory=True) for _ in range(self.nslabs)]
        self.slabs = pool
        self.ready = [threading.Event() for _ in pieces]
        self.events = [None] * len(pieces)
        an.copy_stream.wait_stream(cur)
        self.thread = threading.Thread(target=self._run, args=(an, list(pieces)), daemon=True)
        self.thread.start()

    def _run(self, an, pieces) -> None:
        from concurrent.futures import ThreadPoolExecutor
        try:
            with torch.cuda.device(an.device), ThreadPoolExecutor(self.threads) as pool:
                slab_ev = [None] * self.nslabs
                for i, (a, b) in enumerate(pieces):
                    j = i % self.nslabs
                    if slab_ev[j] is not None:
                        slab_ev[j].synchronize()               # the slab's previous copy has left host memory
                    if b > a:
                        dst = self.slabs[j].numpy()
                        cuts = np.linspace(a, b, self.threads + 1).astype(np.int64)
                        futs = [pool.submit(self.reader.read_into, dst[int(c0) - a:], int(c0), int(c1))
                                for c0, c1 in zip(cuts[:-1], cuts[1:]) if c1 > c0]
                        for f in futs:
                            f.result()
                    with torch.cuda.stream(an.copy_stream):
                        if b > a:
                            an.raw_dev[a:b].copy_(self.slabs[j][:b - a], non_blocking=True)
                        ev = torch.cuda.Event()
                        ev.record(an.copy_stream)
                    slab_ev[j] = self.events[i] = ev
                    self.ready[i].set()
        except BaseException as e:                              # surfaces in wait()
            self.error = e
            for r in self.ready:
                r.set()

    def wait(self, i: int, cur) -> None:
        self.ready[i].wait()
        if self.error is not None:
            raise self.error
        cur.wait_event(self.events[i])

    def finish(self) -> None:
        self.thread.join()

    def whole(self, an: "StreamingAnalyzer") -> torch.Tensor:
        host = torch.empty(an.n_ext, dtype=an.raw_dev.dtype, pin_memory=True)
        self.reader.read_into(host.numpy(), 0, an.n_ext)
        return host


class StreamingAnalyzer:
    """Stages 1-3 over a trace that is still in pinned host memory, overlapped with its own
    host->device copy: the owned range is cut into time sub-shards (the multi-GPU sharding of
    SURVEY.md 8e applied in sequence on one GPU: halo of `required_halo` samples on each side,
    events owned by the sub-shard that holds their start), the copy is cut at the points where
    a sub-shard becomes complete, and each sub-shard is filtered, detected, segmented and its
    tables are sent back while the next pieces are still arriving.  When the last byte lands only
    the last sub-shard is left to do.

    The exact median (the reference's pad value, plot-trace.py:319) is known only at the end; the
    filter needs it only as the pad at the two true ends of the trace (`filtfilt(code - c) + c`
    does not depend on c otherwise), so every sub-shard subtracts the estimate from the first
    piece, the last one is padded with the exact value, and the (small) first one is redone if the
    estimate turns out to be off.  Interior sub-shard boundaries differ from a whole-trace run only by
    the IIR warm-up error (< 1e-7 of the signal range, as between GPUs).  Blocks without enough
    baseline samples inherit the nearest earlier valid block of their own sub-shard."""

    def __init__(self, n_ext: int, settings, cutoff: float, order: int = 8, *, lo_halo: int = 0, hi_halo: int = 0,
                 shards: int = 16, first_blocks: int = 4, group=None, device="cuda", **kw):
        self.n_ext, self.lo_halo, self.hi_halo = int(n_ext), int(lo_halo), int(hi_halo)
        self.n_own = self.n_ext - self.lo_halo - self.hi_halo
        self.settings, self.cutoff, self.order = settings, float(cutoff), int(order)
        self.group, self.device, self.kw = group, torch.device(device), dict(kw)
        self.block = int(kw.get("baseline_block", detect.DEFAULT_BASELINE_BLOCK))
        self.padding = int(kw.get("padding", 1000))
        self.mask = filters.chimera_bitmask(settings)
        self.max_levels = int(kw.get("max_levels", cusum.DEFAULT_MAX_LEVELS))
        self.with_levels = kw.get("cusum_delta") is not None
        fs = float(np.floor(np.squeeze(settings["ADCSAMPLERATE"])))
        max_event = int(kw.get("maxpoints", 100_000)) + 2 * int(kw.get("event_padding", 100))
        self.h = required_halo(self.cutoff, self.order, fs, max_event, self.padding, block=self.block)
        if self.lo_halo % self.block:
            raise ValueError("lo_halo must be a multiple of the baseline block (pipeline.required_halo(..., block=))")
        a0, b_end = self.lo_halo, self.lo_halo + self.n_own
        cuts = [a0]
        if first_blocks * self.block < self.n_own and shards > 1:
            cuts.append(a0 + first_blocks * self.block)
        per = max(self.block, -(-(b_end - cuts[-1]) // max(1, int(shards)) // self.block) * self.block)
        while cuts[-1] < b_end:
            cuts.append(min(b_end, cuts[-1] + per))
        # (a, b, ea, eb): owned [a, b) and extended [ea, eb) ranges of each sub-shard, in raw_ext coordinates
        self.sub = [(a, b, max(0, a - self.h), min(self.n_ext, b + self.h)) for a, b in zip(cuts[:-1], cuts[1:])]
        self.analyzers: dict = {}
        self.y = torch.empty(self.n_own, dtype=torch.float32, device=self.device)
        self.raw_dev = None
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.bmin, self.bmax = float(kw["baseline_min"]), float(kw["baseline_max"])
        self._pin, self._pin_cap = {}, 0
        self._whole = None

    # -- helpers -------------------------------------------------------------------
    def _analyzer(self, a, b, ea, eb) -> TraceAnalyzer:
        key = (eb - ea, a - ea, eb - b)
        if key not in self.analyzers:
            self.analyzers[key] = TraceAnalyzer(eb - ea, self.settings, self.cutoff, self.order, lo_halo=a - ea,
                                                hi_halo=eb - b, group=None, device=self.device,
                                                halos_clipped_by_trace_ends=True, **self.kw)
        return self.analyzers[key]

    def _pinned(self, name: str, like: torch.Tensor, rows: int) -> torch.Tensor:
        """Row-capacity-managed pinned host column (grown geometrically, contents kept)."""
        t = self._pin.get(name)
        if t is None or t.shape[0] < rows:
            cap = max(rows, 2 * (t.shape[0] if t is not None else 0), 4096, self.n_own // 2048)
            new = torch.empty((cap,) + tuple(like.shape[1:]), dtype=like.dtype, pin_memory=True)
            if t is not None:
                torch.cuda.current_stream(self.device).synchronize()
                new[:t.shape[0]] = t
            self._pin[name] = t = new
        return t

    def _process(self, i: int, plan: MedianPlan, pad_x: float, med: tuple[int, int], row0: int, bl: detect.Baseline,
                 end_at: int | None = None) -> int:
        """Analyse sub-shard i (its data must be resident), store its owned samples, blocks and table rows
        (from row `row0`, or ending at row `end_at`); returns its event count (-1: more rows than `end_at`)."""
        a, b, ea, eb = self.sub[i]
        an = self._analyzer(a, b, ea, eb)
        r = an.run(self.raw_dev[ea:eb], _fixed=(plan, pad_x, med))
        a0 = self.lo_halo
        self.y[a - a0:b - a0].copy_(r.filtered)
        k0, nbk, g0 = (a - ea) // self.block, -(-(b - a) // self.block), (a - a0) // self.block
        for name in ("cnt", "s1", "s2", "mean", "std", "sign", "t_start", "t_end"):
            bl.dev[name][g0:g0 + nbk].copy_(r.baseline.dev[name][k0:k0 + nbk])
        nk = int(r.events.starts.numel())
        cols = {"starts": r.events.starts + (a - a0), "ends": r.events.ends + (a - a0), "types": r.types}
        if r.levels is not None:
            cols.update(n_levels=r.levels.n_levels, edges=r.levels.edges, mean=r.levels.mean, std=r.levels.std,
                        overflow=r.levels.overflow)
        if r.intra is not None:
            cols.update(intra_count=r.intra[0], intra_pairs=r.intra[1])
        if end_at is not None:
            if nk > end_at:
                return -1
            row0 = end_at - nk
        for name, t in cols.items():
            self._pinned(name, t, row0 + nk)[row0:row0 + nk].copy_(t, non_blocking=True)
        return nk

    def run_from_host(self, host_codes: torch.Tensor) -> StreamResult:
        """`host_codes`: CPU (ideally pinned) uint16/int16 tensor [lo_halo | owned | hi_halo].  One
        host synchronisation per sub-shard (its event count), all hidden under the copy except the last."""
        if host_codes.numel() != self.n_ext or host_codes.dtype not in (torch.uint16, torch.int16) or host_codes.is_cuda:
            raise ValueError("host_codes must be a CPU uint16/int16 tensor of the planned length")
        with torch.cuda.device(self.device):
            return self._run_stream(_HostFeed(host_codes), host_codes.dtype)

    def run_from_file(self, reader, threads: int = 8, slabs: int = 3) -> StreamResult:
        """The same with the codes still in the `.log` files: `reader` = loader.ChimeraSeries(...).reader(start_s,
        end_s) over [lo_halo | owned | hi_halo].  A worker thread reads the pieces into rotating pinned slabs and
        queues their host->device copies; file reads, PCIe and kernels overlap (the from-file path of SURVEY.md 8d/f4)."""
        if int(reader.n) != self.n_ext:
            raise ValueError(f"analyzer was planned for {self.n_ext} samples, the window holds {reader.n}")
        with torch.cuda.device(self.device):
            return self._run_stream(_FileFeed(reader, threads, slabs), torch.uint16)

    def _run_stream(self, feed, dtype) -> StreamResult:
        if self.raw_dev is None or self.raw_dev.dtype != dtype:
            self.raw_dev = torch.empty(self.n_ext, dtype=dtype, device=self.device)
        L = _lib.lib()
        cur = torch.cuda.current_stream(self.device)
        a0, b_end = self.lo_halo, self.lo_halo + self.n_own
        # ---- the copy, cut where each sub-shard's extended range is complete
        ends = [eb for (_, _, _, eb) in self.sub]
        ends[-1] = self.n_ext
        arrivals, prev = [], 0
        for e in ends:
            arrivals.append((prev, max(prev, e)))
            prev = max(prev, e)
        feed.start(self, arrivals, cur)
        try:
            return self._consume(feed, arrivals, L, cur, a0, b_end)
        finally:
            feed.finish()

    def _consume(self, feed, arrivals, L, cur, a0, b_end) -> StreamResult:
        # ---- median estimate from the first piece
        feed.wait(0, cur)
        first = self.raw_dev[a0:max(a0 + 1, min(arrivals[0][1], b_end))]
        plan = median_estimate(self.n_own, self.mask, _median_kernels(first, self.mask)[0], self.group, self.device,
                               n_sampled=first.numel())
        self.last_plan = plan
        counts = torch.zeros(9, dtype=torch.int64, device=self.device)
        hist_fn, count_fn = _median_kernels(self.raw_dev[a0:b_end], self.mask)
        bl = detect.new_baseline(self.n_own, self.block, self.bmin, self.bmax, self.device)
        st = filters._stream_ptr(self.raw_dev)
        # The first sub-shard holds the true start of the trace, whose pad needs the exact median: it is analysed
        # with the estimate right away and again at the end if the estimate was off.  Its rows sit RIGHT-ALIGNED
        # in a reserved head of the tables, so a different event count the second time moves nothing else.
        head = self.lo_halo == 0 and len(self.sub) > 1
        reserve = max(1024, (self.sub[0][3] - self.sub[0][2]) // 512) if head else 0
        rows, first_rows = reserve, 0
        med, pad_x, redone = (0, 0), 0.0, ""
        # A sub-shard without a single valid baseline block (or more events at the very start than the reserved head
        # holds) sends this rank to the whole-trace analyzer.  That decision must be COLLECTIVE: the rank keeps taking
        # part in the median collectives below, the flag rides with the event counts, and all ranks fall back together.
        failed = False

        def attempt(*args, **kwargs) -> int:
            nonlocal failed
            try:
                return self._process(*args, **kwargs)
            except ValueError as e:
                if "baseline block" not in str(e):
                    raise
                failed = True
                return 0

        for i, (pa, pe) in enumerate(arrivals):
            feed.wait(i, cur)
            ca, cb = max(pa, a0), min(pe, b_end)
            if plan.exact is None and cb > ca:           # exact-median window count of this piece's owned codes
                rc = L.ct_count_window_u16(self.raw_dev[ca:cb].data_ptr(), cb - ca, self.mask, plan.lo, plan.step,
                                           counts.data_ptr(), st)
                _lib.check(rc, "ct_count_window_u16")
            last = i == len(arrivals) - 1
            if last:                                      # everything has arrived: exact order statistics
                med = median_search(self.n_own, self.mask, hist_fn, count_fn, self.group, self.device, plan=plan,
                                    first_counts=counts if plan.exact is None else None)
                pad_x = 0.5 * (med[0] + med[1]) - plan.est
            if failed:
                continue
            if head and i == 0:
                first_rows = attempt(0, plan, 0.0, med, 0, bl, end_at=reserve)
            else:
                rows += attempt(i, plan, pad_x if last else 0.0, med, rows, bl)
        if head and not failed and (pad_x != 0.0 or first_rows < 0):
            first_rows = attempt(0, plan, pad_x, med, 0, bl, end_at=reserve)
            redone = "first"
        if first_rows < 0:                                # more events at the very start than the reserved head holds
            failed = True
        start = reserve - max(first_rows, 0)
        first_id, total, any_failed = event_ids_and_status(0 if failed else rows - start, int(failed), self.group, self.device)
        if any_failed:
            return self._run_whole(feed.whole(self))
        cur.synchronize()
        bl.dev["have_thresholds"] = True
        bl._checked = True
        bl.threshold, bl.hysteresis = float(self.kw.get("threshold", 5.0)), float(self.kw.get("hysteresis", 1.0))
        pad_value = float(np.median(filters.scale_codes_host(np.array(med, dtype=np.uint16), self.settings)))
        tables = {k: v[start:rows].numpy() for k, v in self._pin.items()}
        return StreamResult(filtered=self.y, baseline=bl, tables=tables, pad_value=pad_value, median_codes=tuple(med),
                            first_event_id=first_id, total_events=total, redone=redone)

    def _run_whole(self, host_codes: torch.Tensor) -> StreamResult:
        """Fallback: the whole range through one TraceAnalyzer (baseline inheritance across the whole trace)."""
        if self._whole is None:
            self._whole = TraceAnalyzer(self.n_ext, self.settings, self.cutoff, self.order, lo_halo=self.lo_halo,
                                        hi_halo=self.hi_halo, group=self.group, device=self.device, **self.kw)
        r = self._whole.run_from_host(host_codes)
        t = self._whole.tables_to_host(r)
        return StreamResult(filtered=r.filtered, baseline=r.baseline, tables=t, pad_value=r.pad_value,
                            median_codes=r.median_codes, first_event_id=r.first_event_id, total_events=r.total_events,
                            redone="whole")


def analyze_shard(raw_ext: torch.Tensor, settings, cutoff: float, order: int, *, lo_halo: int, hi_halo: int,
                  threshold: float, hysteresis: float, baseline_block: int, baseline_min: float,
                  baseline_max: float, group=None, is_first: bool = True, is_last: bool = True,
                  padding: int = 1000, keep_filtered: bool = True) -> TraceResult:
    """Run stages 1-2 on a time shard (see TraceAnalyzer; this is the one-shot form)."""
    an = TraceAnalyzer(raw_ext.numel(), settings, cutoff, order, lo_halo=lo_halo, hi_halo=hi_halo, threshold=threshold,
                       hysteresis=hysteresis, baseline_block=baseline_block, baseline_min=baseline_min,
                       baseline_max=baseline_max, padding=padding, group=group, device=raw_ext.device)
    r = an.run(raw_ext)
    return TraceResult(filtered=r.filtered if keep_filtered else r.filtered[:0], baseline=r.baseline, events=r.events,
                       pad_value=r.pad_value, median_codes=r.median_codes, first_event_id=r.first_event_id,
                       total_events=r.total_events)


def analyze_trace(raw: torch.Tensor, settings, cutoff: float, order: int = 8, *, threshold: float = 5.0,
                  hysteresis: float = 1.0, baseline_block: int = detect.DEFAULT_BASELINE_BLOCK,
                  baseline_min: float, baseline_max: float, padding: int = 1000) -> TraceResult:
    """Single-GPU stages 1-2 over a whole device-resident trace."""
    return analyze_shard(raw, settings, cutoff, order, lo_halo=0, hi_halo=0, threshold=threshold,
                         hysteresis=hysteresis, baseline_block=baseline_block, baseline_min=baseline_min,
                         baseline_max=baseline_max, padding=padding)


def gather_tables(tables: dict, group=None, dst: int = 0) -> dict | None:
    """Gather per-rank event-table columns (tensors whose first dimension is the rank's event count) to rank
    `dst` in rank order = time order (the one data-sized collective of the path, O(events)).
    One small all_reduce exchanges the row counts; every column then travels point to point, exactly its rows,
    straight into its slice of the concatenated column on `dst` (one batched send/receive group, no padding and
    nothing sent to ranks that do not need it).  Returns the concatenated columns on `dst`, None elsewhere; with
    `group=None` the input itself."""
    if group is None:
        return tables
    import torch.distributed as dist
    ws, rk = dist.get_world_size(group), dist.get_rank(group)
    first = next(iter(tables.values()))
    dev = first.device
    counts = torch.zeros(ws, dtype=torch.int64, device=dev)
    counts[rk] = first.shape[0]
    counts = _all_reduce_(counts, group).cpu().numpy()
    offs = np.concatenate(([0], np.cumsum(counts)))
    out, ops = {}, []
    for name, t in tables.items():                       # same column order on every rank (dict order)
        t = t.contiguous()
        if rk == dst:
            full = torch.empty((int(offs[-1]),) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
            full[int(offs[rk]):int(offs[rk + 1])] = t
            for src in range(ws):
                if src != dst and counts[src]:
                    ops.append(dist.P2POp(dist.irecv, full[int(offs[src]):int(offs[src + 1])], dist.get_global_rank(group, src), group))
            out[name] = full
        elif counts[rk]:
            ops.append(dist.P2POp(dist.isend, t, dist.get_global_rank(group, dst), group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return out if rk == dst else None


def shard_bounds(n: int, world: int, rank: int, align: int) -> tuple[int, int]:
    """Owned range of `rank`: equal shares rounded to `align` (the baseline block), so
    baseline blocks never straddle ranks."""
    per = -(-n // world)
    per = -(-per // align) * align
    return min(n, rank * per), min(n, (rank + 1) * per)
