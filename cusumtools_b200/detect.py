"""Stage 2 of the hot path: baseline block statistics and threshold/hysteresis event
detection with stream compaction, on the GPU.

The reference contains no implementation of this stage; it draws and consumes its output
(`plot-trace.py:172-203,350-414`: baseline.csv rows `time_s, baseline_pA, stdev_pA`, start
line `baseline - sign*threshold*stdev`, end line `baseline - sign*(threshold-hysteresis)*
stdev`; rate.csv rows `id, type, start_time_s, end_time_s`).  Definition of record:
oracle/events_oracle.py.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib
from .filters import _require_cuda, _stream_ptr

STATS_MAX_SHIFT = 8
DEFAULT_BASELINE_BLOCK = 1 << 20


def stats_shift(half_width: float, block: int) -> int:
    """Fixed-point fraction bits of the block sums: largest s <= 8 with
    (half_width*2^s + 1)^2 * block < 2^62."""
    s = STATS_MAX_SHIFT
    while s > -32 and (float(half_width) * 2.0 ** s + 1.0) ** 2 * float(block) >= 2.0 ** 62:
        s -= 1
    return s


@dataclass
class Baseline:
    """Per-block baseline table (the rows of baseline.csv) plus the float32 lines the
    detector compares against."""
    block: int
    mean: np.ndarray      # float64 [nb]  baseline_pA
    std: np.ndarray       # float64 [nb]  stdev_pA
    count: np.ndarray     # int64   [nb]  samples inside the baseline window
    sign: np.ndarray | None = None      # int32 [nb]
    t_start: np.ndarray | None = None   # float32 [nb]
    t_end: np.ndarray | None = None     # float32 [nb]

    def with_thresholds(self, threshold: float, hysteresis: float) -> "Baseline":
        sign = np.where(self.mean >= 0, 1, -1).astype(np.int32)
        self.sign = sign
        self.t_start = (self.mean - sign * threshold * self.std).astype(np.float32)
        self.t_end = (self.mean - sign * (threshold - hysteresis) * self.std).astype(np.float32)
        return self


def baseline_blocks(y: torch.Tensor, block: int, baseline_min: float, baseline_max: float,
                    min_count: int = 16) -> Baseline:
    """Mean / population std of the samples inside [baseline_min, baseline_max] for every
    block of `block` samples.  The device produces exact integer sums; the (tiny) division
    and square root run on the host in Python integers / float64."""
    _require_cuda(y, "y", torch.float32)
    L = _lib.lib()
    n = y.numel()
    block = int(block)
    nb = (n + block - 1) // block
    c0 = np.float32(0.5 * (np.float32(baseline_min) + np.float32(baseline_max)))
    hw = max(float(c0) - float(np.float32(baseline_min)), float(np.float32(baseline_max)) - float(c0))
    shift = stats_shift(hw, block)
    acc = torch.zeros((3, max(nb, 1)), dtype=torch.int64, device=y.device)
    rc = L.ct_block_stats_f32(y.data_ptr(), n, block, float(baseline_min), float(baseline_max), float(c0), shift,
                              acc[0].data_ptr(), acc[1].data_ptr(), acc[2].data_ptr(), _stream_ptr(y))
    _lib.check(rc, "ct_block_stats_f32")
    cnt, s1, s2 = acc.cpu().numpy()
    mean = np.full(nb, np.nan)
    std = np.full(nb, np.nan)
    sc = 2.0 ** shift
    for k in range(nb):
        c = int(cnt[k])
        if c >= min_count:
            a, b = int(s1[k]), int(s2[k])
            mean[k] = float(c0) + (a / c) / sc
            std[k] = math.sqrt(max(b * c - a * a, 0) / (c * c)) / sc
    valid = np.nonzero(~np.isnan(mean))[0]
    if nb and valid.size == 0:
        raise ValueError("no baseline block has enough samples inside [baseline_min, baseline_max]")
    if nb:
        last = valid[0]
        for k in range(nb):
            if np.isnan(mean[k]):
                mean[k], std[k] = mean[last], std[last]
            else:
                last = k
    return Baseline(block=block, mean=mean, std=std, count=cnt[:nb].copy())


@dataclass
class EventList:
    starts: torch.Tensor      # int64 [E] device, first sample beyond the start line
    ends: torch.Tensor        # int64 [E] device, first sample back beyond the end line
    open_start: int           # start of an event still open at the end of the data, or -1

    def __len__(self) -> int:
        return int(self.starts.numel())


def detect_events(y: torch.Tensor, baseline: Baseline, *, state_in: bool = False,
                  capacity: int | None = None) -> EventList:
    """Threshold/hysteresis detection over the whole filtered trace `y` (device float32).
    Event i occupies samples [starts[i], ends[i])."""
    _require_cuda(y, "y", torch.float32)
    if baseline.t_start is None:
        raise ValueError("call Baseline.with_thresholds(threshold, hysteresis) first")
    L = _lib.lib()
    n = y.numel()
    run = L.ct_detect_run()
    if baseline.block % run:
        raise ValueError(f"baseline block must be a multiple of {run} samples")
    dev = y.device
    sign = torch.from_numpy(np.ascontiguousarray(baseline.sign, np.int32)).to(dev)
    ts = torch.from_numpy(np.ascontiguousarray(baseline.t_start, np.float32)).to(dev)
    te = torch.from_numpy(np.ascontiguousarray(baseline.t_end, np.float32)).to(dev)
    wsb = int(L.ct_detect_workspace_bytes(n))
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    cap = int(capacity) if capacity else max(1024, n // 2048)
    while True:
        starts = torch.empty(cap, dtype=torch.int64, device=dev)
        ends = torch.empty(cap, dtype=torch.int64, device=dev)
        counts = torch.zeros(2, dtype=torch.int64, device=dev)
        rc = L.ct_detect_f32(y.data_ptr(), n, baseline.block, sign.data_ptr(), ts.data_ptr(), te.data_ptr(),
                             int(bool(state_in)), ws.data_ptr(), wsb, starts.data_ptr(), ends.data_ptr(), cap,
                             counts.data_ptr(), _stream_ptr(y))
        _lib.check(rc, "ct_detect_f32")
        ns, ne = (int(v) for v in counts.cpu().numpy())
        if max(ns, ne) <= cap:
            break
        cap = max(ns, ne)
    starts, ends = starts[:ns], ends[:ne]
    if state_in and ne and (ns == 0 or ne > ns or int(ends[0]) < int(starts[0])):
        ends = ends[1:]          # closes an event opened before this data
        ne -= 1
    open_start = -1
    if ns > ne:
        open_start = int(starts[ne])
        starts = starts[:ne]
    return EventList(starts=starts, ends=ends, open_start=open_start)


def event_windows(starts: torch.Tensor, ends: torch.Tensor, n: int, padding: int, minpoints: int,
                  maxpoints: int):
    """Sample windows [start - padding, end + padding) handed to CUSUM+ and the rate.csv
    `type` code: 0 accepted, 2 too short, 3 too long, 4 padding leaves the trace or overlaps
    a neighbouring event (plot-trace.py:354-357 treats type > 1 as rejected)."""
    w0 = starts - padding
    w1 = ends + padding
    length = ends - starts
    z = torch.zeros(1, dtype=torch.int64, device=starts.device)
    nn = torch.full((1,), n, dtype=torch.int64, device=starts.device)
    prev_end = torch.cat((z, ends[:-1]))
    next_start = torch.cat((starts[1:], nn))
    typ = torch.zeros_like(starts, dtype=torch.int32)
    typ[(w0 < prev_end) | (w1 > next_start) | (w0 < 0) | (w1 > n)] = 4
    typ[length > maxpoints] = 3
    typ[length < minpoints] = 2
    return w0, w1, typ
