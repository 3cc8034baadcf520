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


def _all_reduce_(t: torch.Tensor, group) -> torch.Tensor:
    if group is not None:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def median_search(n_local: int, mask: int, hist_fn, count_fn, group=None, device=None) -> tuple[int, int]:
    """Host side of the exact global median: the two middle order statistics of the masked
    codes of ALL ranks.  `hist_fn(stride)` returns this rank's strided-sample histogram
    (int tensor[65536]) and `count_fn(lo, step)` its 9 window counters (#codes < lo, then
    #codes == lo + i*step); both are summed over `group` here, so every rank takes the same
    decisions and returns the same pair.  (The device kernels are passed in, which is what
    lets the world-size-2 gloo test drive this logic on CPU.)"""
    nt = torch.tensor([int(n_local)], dtype=torch.int64, device=device)
    n = int(_all_reduce_(nt, group).item())
    if n == 0:
        raise ValueError("median of an empty trace")
    shift = 0
    while shift < 16 and not (mask >> shift) & 1:
        shift += 1
    step = 1 << shift
    k1, k2 = (n - 1) // 2, n // 2
    stride = max(1, n // (1 << 22))
    h = _all_reduce_(hist_fn(stride).to(torch.int64), group)
    cdf = np.cumsum(h.cpu().numpy())
    if stride == 1:
        return int(np.searchsorted(cdf, k1 + 1)), int(np.searchsorted(cdf, k2 + 1))
    est = int(np.searchsorted(cdf, (cdf[-1] + 1) // 2))
    lo = max(0, (est >> shift) * step - 3 * step)
    for _ in range(16):
        c = _all_reduce_(count_fn(lo, step).to(torch.int64), group).cpu().numpy().astype(np.int64)
        below, cw = int(c[0]), int(c[0]) + np.cumsum(c[1:])
        if below <= k1 and k2 < cw[-1]:
            return lo + int(np.searchsorted(cw, k1 + 1)) * step, lo + int(np.searchsorted(cw, k2 + 1)) * step
        lo = max(0, lo - 6 * step) if k1 < below else lo + 6 * step
    cdf = np.cumsum(_all_reduce_(hist_fn(1).to(torch.int64), group).cpu().numpy())   # pathological distribution
    return int(np.searchsorted(cdf, k1 + 1)), int(np.searchsorted(cdf, k2 + 1))


def global_code_median(raw_owned: torch.Tensor, mask: int, group=None) -> tuple[int, int]:
    """Exact median (two middle order statistics) of the masked codes of the WHOLE trace
    when each rank passes its owned samples; identical to filters.code_median for one rank."""
    if group is None:
        return filters.code_median(raw_owned, mask)
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

    return median_search(n_local, mask, hist_fn, count_fn, group, dev)


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


def required_halo(cutoff: float, order: int, samplerate: float, max_event: int, padding: int = 1000,
                  eps: float = filters.DEFAULT_HALO_EPS) -> int:
    """Samples a rank must read beyond each end of its owned range: IIR warm-up on both
    sides (forward and backward pass) plus one maximal event so that an event straddling
    a shard boundary is seen whole by the rank that owns its start."""
    d = bessel_lowpass(int(order), 2.0 * float(cutoff) / float(samplerate))
    return 2 * filters.halo_samples(d, eps) + int(max_event)


@dataclass
class AnalysisResult:
    """Everything one pass of the hot path over a (shard of a) trace produces; all tensors
    are device-resident views into the analyzer's buffers (valid until its next run)."""
    filtered: torch.Tensor          # float32, the rank's owned samples
    detect_trace: torch.Tensor      # float32, owned + right halo: what event indices refer to
    baseline: detect.Baseline
    events: detect.EventList        # indices relative to the rank's first owned sample
    win_start: torch.Tensor         # int64 [E]  CUSUM+ windows [start - padding, end + padding)
    win_end: torch.Tensor           # int64 [E]
    types: torch.Tensor             # int32 [E]  rate.csv type code (0 accepted, >1 rejected)
    levels: "cusum.LevelTable | None"
    pad_value: float
    median_codes: tuple[int, int]
    first_event_id: int = 0
    total_events: int = 0


class TraceAnalyzer:
    """Stages 1-3 over a time shard with every buffer allocated once and a single host
    synchronisation at the end of the step (the median needs two small ones up front).

    `raw_ext` = [lo_halo | owned | hi_halo] codes.  The filter runs over the extended range
    (its constant pad only matters at the true ends of the trace); detection and CUSUM+ run
    over [owned | hi_halo] so an event that starts in the owned range and ends in the halo
    is completed, and events starting in the halo are left to the next rank."""

    def __init__(self, n_ext: int, settings, cutoff: float, order: int = 8, *, lo_halo: int = 0, hi_halo: int = 0,
                 threshold: float = 5.0, hysteresis: float = 1.0, baseline_block: int = detect.DEFAULT_BASELINE_BLOCK,
                 baseline_min: float, baseline_max: float, padding: int = 1000, event_padding: int = 100,
                 minpoints: int = 8, maxpoints: int = 100_000, cusum_delta: float | None = None,
                 cusum_h: float | None = None, max_levels: int = cusum.DEFAULT_MAX_LEVELS,
                 event_capacity: int | None = None, group=None, device="cuda"):
        self.n_ext, self.lo_halo, self.hi_halo = int(n_ext), int(lo_halo), int(hi_halo)
        self.n_own = self.n_ext - self.lo_halo - self.hi_halo
        self.n_det = self.n_ext - self.lo_halo
        self.settings, self.cutoff, self.order, self.padding = settings, float(cutoff), int(order), int(padding)
        self.threshold, self.hysteresis = float(threshold), float(hysteresis)
        self.block, self.bmin, self.bmax = int(baseline_block), float(baseline_min), float(baseline_max)
        self.event_padding, self.minpoints, self.maxpoints = int(event_padding), int(minpoints), int(maxpoints)
        self.delta, self.h, self.max_levels = cusum_delta, cusum_h, int(max_levels)
        self.group = group
        self.device = torch.device(device)
        self.mask = filters.chimera_bitmask(settings)
        L = _lib.lib()
        if self.block % L.ct_detect_run():
            raise ValueError(f"baseline block must be a multiple of {L.ct_detect_run()} samples")
        self.y = torch.empty(self.n_ext, dtype=torch.float32, device=self.device)
        self.design = bessel_lowpass(self.order, 2.0 * self.cutoff / float(np.floor(np.squeeze(settings["ADCSAMPLERATE"]))))
        # baseline block sums ride on the filter's epilogue when the block is a whole number of its warp groups
        self.fuse_stats = self.block % filters.stats_granule(self.n_ext, self.padding, self.design) == 0
        self.filter_ws = None
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
        self.scalars = torch.zeros(4, dtype=torch.int64, device=dev)     # n_starts, n_ends, n_kept
        if self.delta is not None:
            self.nl = torch.empty(cap, dtype=torch.int32, device=dev)
            self.ed = torch.empty((cap, ML + 1), dtype=torch.int32, device=dev)
            self.mu = torch.empty((cap, ML), dtype=torch.float64, device=dev)
            self.sd = torch.empty((cap, ML), dtype=torch.float64, device=dev)
            self.ov = torch.empty(cap, dtype=torch.uint8, device=dev)
            self.cws_bytes = int(_lib.lib().ct_cusum_workspace_bytes(cap))
            self.cws = torch.empty((self.cws_bytes + 7) // 8, dtype=torch.int64, device=dev)

    def run(self, raw_ext: torch.Tensor, stage_hook=None) -> AnalysisResult:
        """One pass of stages 1-3.  `stage_hook(name)` (optional) is called after the launches of each
        stage have been enqueued: 'median', 'filter', 'baseline', 'detect', 'cusum' (profiling only)."""
        hook = stage_hook or (lambda name: None)
        if raw_ext.numel() != self.n_ext:
            raise ValueError(f"analyzer was planned for {self.n_ext} samples, got {raw_ext.numel()}")
        L = _lib.lib()
        lo, n_own = self.lo_halo, self.n_own
        owned = raw_ext[lo:lo + n_own]
        c1, c2 = global_code_median(owned, self.mask, self.group)
        hook("median")
        if self.filter_ws is None:
            need = int(L.ct_filtfilt_workspace_bytes(self.n_ext, self.padding, max(1, self.design.impulse_tail(filters.DEFAULT_HALO_EPS))))
            self.filter_ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        bl = detect.new_baseline(self.n_det, self.block, self.bmin, self.bmax, self.device) if self.fuse_stats else None
        y = filters.dequant_filtfilt(raw_ext, self.settings, self.cutoff, self.order, padding=self.padding,
                                     median_codes=(c1, c2), out=self.y, workspace=self.filter_ws,
                                     stats=detect.stats_args(bl, origin=lo) if bl is not None else None)
        pad_value = float(np.median(filters.scale_codes_host(np.array([c1, c2], dtype=np.uint16), self.settings)))
        yd = y[lo:]
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
                                 self.ws.data_ptr(), self.ws_bytes, self.starts.data_ptr(), self.ends.data_ptr(),
                                 self.cap, sc[0:].data_ptr(), st)
            _lib.check(rc, "ct_detect_f32")
            rc = L.ct_event_windows(self.starts.data_ptr(), self.ends.data_ptr(), sc[0:].data_ptr(), self.cap, self.n_det,
                                    n_own, self.event_padding, self.minpoints, self.maxpoints, self.w0.data_ptr(),
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
            hook("cusum")
            host = torch.cat((sc[:3], bl.dev["status"].to(torch.int64))).cpu().numpy()   # the step's one sync
            ns, ne, nk = int(host[0]), int(host[1]), int(host[2])
            if max(ns, ne) <= self.cap:
                break
            self._alloc_events(max(ns, ne))          # more events than planned: grow and redo stages 2-3
        if int(host[3]) != 0:
            raise ValueError("no baseline block has enough samples inside [baseline_min, baseline_max]")
        bl._checked = True
        open_start = -1
        if ns > ne:                                   # an event still open at the end of the data
            o = int(self.starts[ne].item())
            open_start = o if o < n_own else -1
        ev = detect.EventList(self.starts[:nk], self.ends[:nk], open_start)
        lv = None
        if self.delta is not None:
            lv = cusum.LevelTable(self.nl[:nk], self.ed[:nk], self.mu[:nk], self.sd[:nk], self.ov[:nk], self.max_levels)
        first_id, total = event_id_offsets(nk, self.group, self.device)
        return AnalysisResult(filtered=y[lo:lo + n_own], detect_trace=yd, baseline=bl, events=ev,
                              win_start=self.w0[:nk], win_end=self.w1[:nk], types=self.typ[:nk], levels=lv,
                              pad_value=pad_value, median_codes=(c1, c2), first_event_id=first_id, total_events=total)


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
        out = {}
        for k, t in src.items():
            if k not in self._pinned:
                self._pinned[k] = torch.empty((self.cap,) + tuple(t.shape[1:]), dtype=t.dtype, pin_memory=True)
            h = self._pinned[k][:nk]
            h.copy_(t, non_blocking=True)
            out[k] = h
        torch.cuda.current_stream(self.device).synchronize()
        return {k: v.numpy() for k, v in out.items()}


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


def shard_bounds(n: int, world: int, rank: int, align: int) -> tuple[int, int]:
    """Owned range of `rank`: equal shares rounded to `align` (the baseline block), so
    baseline blocks never straddle ranks."""
    per = -(-n // world)
    per = -(-per // align) * align
    return min(n, rank * per), min(n, (rank + 1) * per)
