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

from . import _lib, detect, filters
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


def global_code_median(raw_owned: torch.Tensor, mask: int, group=None) -> tuple[int, int]:
    """Exact median (two middle order statistics) of the masked codes of the WHOLE trace
    when each rank passes its owned samples; identical to filters.code_median for one rank."""
    if group is None:
        return filters.code_median(raw_owned, mask)
    import torch.distributed as dist
    L = _lib.lib()
    dev = raw_owned.device
    n_local = raw_owned.numel()
    nt = torch.tensor([n_local], dtype=torch.int64, device=dev)
    _all_reduce_(nt, group)
    n = int(nt.item())
    shift = 0
    while shift < 16 and not (mask >> shift) & 1:
        shift += 1
    step = 1 << shift
    k1, k2 = (n - 1) // 2, n // 2
    st = filters._stream_ptr(raw_owned)
    stride = max(1, n // (1 << 22))
    h = torch.zeros(65536, dtype=torch.int32, device=dev)
    _lib.check(L.ct_hist_sampled_u16(raw_owned.data_ptr(), n_local, stride, mask, h.data_ptr(), st), "ct_hist_sampled_u16")
    h = _all_reduce_(h.to(torch.int64), group)
    cdf = np.cumsum(h.cpu().numpy())
    if stride == 1:
        return int(np.searchsorted(cdf, k1 + 1)), int(np.searchsorted(cdf, k2 + 1))
    est = int(np.searchsorted(cdf, (cdf[-1] + 1) // 2))
    lo = max(0, (est >> shift) * step - 3 * step)
    for _ in range(16):
        cnt = torch.zeros(9, dtype=torch.int64, device=dev)
        _lib.check(L.ct_count_window_u16(raw_owned.data_ptr(), n_local, mask, lo, step, cnt.data_ptr(), st), "ct_count_window_u16")
        c = _all_reduce_(cnt, group).cpu().numpy().astype(np.int64)
        below, cw = int(c[0]), int(c[0]) + np.cumsum(c[1:])
        if below <= k1 and k2 < cw[-1]:
            return lo + int(np.searchsorted(cw, k1 + 1)) * step, lo + int(np.searchsorted(cw, k2 + 1)) * step
        lo = max(0, lo - 6 * step) if k1 < below else lo + 6 * step
    raise RuntimeError("median window search did not converge")


def required_halo(cutoff: float, order: int, samplerate: float, max_event: int, padding: int = 1000,
                  eps: float = filters.DEFAULT_HALO_EPS) -> int:
    """Samples a rank must read beyond each end of its owned range: IIR warm-up on both
    sides (forward and backward pass) plus one maximal event so that an event straddling
    a shard boundary is seen whole by the rank that owns its start."""
    d = bessel_lowpass(int(order), 2.0 * float(cutoff) / float(samplerate))
    return 2 * filters.halo_samples(d, eps) + int(max_event)


def analyze_shard(raw_ext: torch.Tensor, settings, cutoff: float, order: int, *, lo_halo: int, hi_halo: int,
                  threshold: float, hysteresis: float, baseline_block: int, baseline_min: float,
                  baseline_max: float, group=None, is_first: bool = True, is_last: bool = True,
                  padding: int = 1000, keep_filtered: bool = True) -> TraceResult:
    """Run stages 1-2 on a time shard.  `raw_ext` = [lo_halo | owned | hi_halo] codes.

    The filter runs over the extended range (its constant pad only matters at the true
    ends of the trace, i.e. on the first/last rank where the halo is 0); detection runs
    over [owned | hi_halo] so an event that starts in the owned range and ends in the halo
    is completed, and events starting in the halo are left to the next rank."""
    n_ext = raw_ext.numel()
    owned = raw_ext[lo_halo:n_ext - hi_halo]
    mask = filters.chimera_bitmask(settings)
    c1, c2 = global_code_median(owned, mask, group)
    y_ext = filters.dequant_filtfilt(raw_ext, settings, cutoff, order, padding=padding, median_codes=(c1, c2))
    pad_value = float(np.median(filters.scale_codes_host(np.array([c1, c2], dtype=np.uint16), settings)))
    n_own = owned.numel()
    y_det = y_ext[lo_halo:]                      # owned + right halo
    bl = detect.baseline_blocks(y_det, baseline_block, baseline_min, baseline_max).with_thresholds(threshold, hysteresis)
    ev = detect.detect_events(y_det, bl)
    keep = ev.starts < n_own
    nk = int(keep.sum().item()) if len(ev) else 0
    open_start = ev.open_start if (ev.open_start >= 0 and ev.open_start < n_own) else -1
    ev = detect.EventList(ev.starts[:nk], ev.ends[:nk], open_start)
    first_id, total = 0, nk
    if group is not None:
        import torch.distributed as dist
        ws, rk = dist.get_world_size(group), dist.get_rank(group)
        counts = torch.zeros(ws, dtype=torch.int64, device=raw_ext.device)
        counts[rk] = nk
        _all_reduce_(counts, group)
        c = counts.cpu().numpy()
        first_id, total = int(c[:rk].sum()), int(c.sum())
    return TraceResult(filtered=y_ext[lo_halo:lo_halo + n_own] if keep_filtered else y_ext[:0], baseline=bl,
                       events=ev, pad_value=pad_value, median_codes=(c1, c2), first_event_id=first_id,
                       total_events=total)


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
