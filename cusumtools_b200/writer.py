"""The event table and the analysis directory the reference's consumers read.

`readevents.py:1527-1544` opens `events.csv` (+ `rate.csv`, `summary.txt`, `events/`) and
`plot-trace.py:172-203` opens `rate.csv`, `baseline.csv`, `summary.txt`; the reference
contains no producer of these files except `mosaicConverter.py:38-194`, which is therefore
the in-repo definition of every column (names, units, `%.16g;` list format).  Conventions
fixed here (SURVEY.md Appendix A.4 lists the open points):

* the CUSUM+ window of an event is [start - padding, end + padding); its first and last
  level are the baseline before / after the event, so the four list columns carry them as
  first and last entry, as the consumers expect (`readevents.py:844,1300-1302`);
* `effective_baseline_pA` = mean of the two; a blockage is `sign(baseline) * (baseline -
  level)`, positive when the current magnitude drops; `blockages_pA` = [before - eff] +
  sub-level blockages + [after - eff] (`mosaicConverter.py:131-132`);
* `n_levels` = number of sub-levels + 1 (`mosaicConverter.py:79`);
* in-event statistics (`average_blockage_pA`, `area_pC`) run over the sub-levels, i.e. from
  the first to the last changepoint; `area_pC` = average blockage (pA) x that duration (s);
* `residual_pA` = rms of trace minus step fit over the whole window = sqrt(sum len*std^2 / sum len);
* `max_deviation_pA` = largest |sample - effective baseline| in the window;
* `type`: 0 accepted (CUSUM fit; 3-column event file), 2 too short, 3 too long, 4 padding
  overlaps a neighbour or leaves the trace, 5 CUSUM+ found no sub-level, 6 more levels than
  the table holds; types > 1 appear in rate.csv only (`plot-trace.py:354-357`);
* intra-event threshold crossings (`intra_crossings`, rate.csv `intra_crossing_times_us`) are
  produced when the analyzer is given `intra_threshold` > 0 (detect.intra_crossings; the lines the
  consumer draws at readevents.py:1363-1366); at most `max_crossings` pairs are listed per event,
  the count is complete.  No step-response fit: `rc_const1_us` = `rc_const2_us` = 0.

Everything here is O(events) host arithmetic on tables that came from the device; the only
device work is the per-event extrema kernel (ct_event_extrema_f32).
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib
from .filters import _require_cuda, _stream_ptr

EVENT_COLUMNS = ["id", "type", "start_time_s", "event_delay_s", "duration_us", "threshold", "baseline_before_pA",
                 "baseline_after_pA", "effective_baseline_pA", "area_pC", "average_blockage_pA",
                 "relative_average_blockage", "max_blockage_pA", "relative_max_blockage", "max_blockage_duration_us",
                 "n_levels", "intra_crossings", "rc_const1_us", "rc_const2_us", "residual_pA", "max_deviation_pA",
                 "min_blockage_pA", "relative_min_blockage", "min_blockage_duration_us", "level_current_pA",
                 "level_duration_us", "blockages_pA", "stdev_pA"]
RATE_COLUMNS = ["id", "type", "start_time_s", "end_time_s", "intra_crossing_times_us", "local_stdev", "local_baseline"]
TYPE_NO_SUBLEVEL, TYPE_LEVEL_OVERFLOW = 5, 6


def event_extrema(y: torch.Tensor, win_start: torch.Tensor, win_end: torch.Tensor):
    """(min, max) sample of every event window, float32 device tensors."""
    _require_cuda(y, "y", torch.float32)
    _require_cuda(win_start, "win_start", torch.int64)
    _require_cuda(win_end, "win_end", torch.int64)
    E = win_start.numel()
    lo = torch.empty(E, dtype=torch.float32, device=y.device)
    hi = torch.empty(E, dtype=torch.float32, device=y.device)
    rc = _lib.lib().ct_event_extrema_f32(y.data_ptr(), y.numel(), win_start.data_ptr(), win_end.data_ptr(), E, None,
                                         lo.data_ptr(), hi.data_ptr(), _stream_ptr(y))
    _lib.check(rc, "ct_event_extrema_f32")
    return lo, hi


@dataclass
class EventTable:
    """Column-oriented event table (numpy).  `rate` holds every detected event, `events`
    the accepted ones (type 0) with the full column set; list columns are object arrays of
    float64 vectors until they are serialised."""
    events: dict
    rate: dict

    def __len__(self) -> int:
        return len(self.events["id"])


def _fmt_list(v) -> str:
    return ";".join("%.16g" % x for x in v)        # mosaicConverter.py:139-148 ('%.16g;' joined, last ';' cut)


def build_event_table(*, starts, ends, types, n_levels, edges, level_mean, level_std, overflow, xmin, xmax,
                      samplerate: float, threshold: float, baseline_mean, baseline_std, baseline_block: int,
                      padding: int, first_id: int = 0, time_offset_s: float = 0.0, index_offset: int = 0,
                      block_offset: int = 0, intra_count=None, intra_pairs=None) -> EventTable:
    """Per-event columns from the detector / CUSUM+ tables (numpy arrays, one row per detected
    event).  `index_offset` is the global sample index of the shard's first owned sample,
    `first_id` the global id of its first event (multi-GPU: pipeline.AnalysisResult)."""
    starts = np.asarray(starts, np.int64); ends = np.asarray(ends, np.int64)
    E = starts.size
    types = np.asarray(types, np.int32).copy()
    nl = np.asarray(n_levels, np.int64)
    edges = np.asarray(edges, np.int64).reshape(E, -1)
    mu = np.asarray(level_mean, np.float64).reshape(E, -1)
    sd = np.asarray(level_std, np.float64).reshape(E, -1)
    ML = mu.shape[1] if E else 0
    fs = float(samplerate)
    types[(types == 0) & (np.asarray(overflow) != 0)] = TYPE_LEVEL_OVERFLOW
    types[(types == 0) & (nl < 3)] = TYPE_NO_SUBLEVEL
    ids = first_id + np.arange(E, dtype=np.int64)
    t_start = time_offset_s + (index_offset + starts) / fs
    t_end = time_offset_s + (index_offset + ends) / fs
    # `block_offset`: samples the baseline table starts before the first owned sample (a shard's left halo)
    blk = np.minimum((starts + int(block_offset)) // int(baseline_block), len(baseline_mean) - 1) if E else np.zeros(0, np.int64)
    # intra-event crossings: ';'-joined start;end;start;end... in us on the event file's time axis (0 = window start)
    crossing_txt = np.array([""] * E, dtype=object)
    ic = np.zeros(E, np.int64)
    if intra_count is not None and E:
        ic = np.asarray(intra_count, np.int64)
        ip = np.asarray(intra_pairs, np.int64).reshape(E, -1)
        kmax = ip.shape[1] // 2
        for i in np.nonzero(ic > 0)[0]:
            crossing_txt[i] = ";".join("%.16g" % (v * (1e6 / fs)) for v in ip[i, :2 * min(int(ic[i]), kmax)])
    rate = {"id": ids, "type": types, "start_time_s": t_start, "end_time_s": t_end,
            "intra_crossing_times_us": crossing_txt,
            "local_stdev": np.asarray(baseline_std, np.float64)[blk] if E else np.zeros(0),
            "local_baseline": np.asarray(baseline_mean, np.float64)[blk] if E else np.zeros(0)}

    ok = np.nonzero(types == 0)[0]
    n = ok.size
    L = nl[ok]
    col = np.arange(ML)[None, :]
    valid = col < L[:, None]                       # levels of the window
    inner = (col >= 1) & (col < (L - 1)[:, None])  # sub-levels (between first and last changepoint)
    length = np.where(valid, edges[ok, 1:ML + 1] - edges[ok, :ML], 0).astype(np.float64)
    m = np.where(valid, mu[ok], 0.0)
    s = np.where(valid, sd[ok], 0.0)
    rows = np.arange(n)
    before = m[rows, 0]
    after = m[rows, np.maximum(L - 1, 0)]
    eff = 0.5 * (before + after)
    sgn = np.where(eff >= 0, 1.0, -1.0)
    block = sgn[:, None] * (eff[:, None] - m)      # blockage of every level
    inner_len = np.where(inner, length, 0.0)
    dur_in = inner_len.sum(axis=1)
    with np.errstate(invalid="ignore", divide="ignore"):
        mean_in = (inner_len * m).sum(axis=1) / dur_in
        avg_block = sgn * (eff - mean_in)
        big = np.where(inner, block, -np.inf); small = np.where(inner, block, np.inf)
        imax = big.argmax(axis=1) if n else np.zeros(0, np.int64)
        imin = small.argmin(axis=1) if n else np.zeros(0, np.int64)
        max_block, min_block = block[rows, imax], block[rows, imin]
        residual = np.sqrt((length * s * s).sum(axis=1) / length.sum(axis=1))
        aeff = np.abs(eff)
        xmin = np.asarray(xmin, np.float64)[ok]; xmax = np.asarray(xmax, np.float64)[ok]
        max_dev = np.maximum(np.abs(xmax - eff), np.abs(xmin - eff))
        us = 1e6 / fs
        lists = {k: np.empty(n, dtype=object) for k in ("level_current_pA", "level_duration_us", "blockages_pA", "stdev_pA")}
        for i in range(n):
            k = int(L[i])
            bl = block[i, :k].copy()
            bl[0] = before[i] - eff[i]; bl[k - 1] = after[i] - eff[i]      # mosaicConverter.py:131-132
            lists["level_current_pA"][i] = m[i, :k].copy()
            lists["level_duration_us"][i] = length[i, :k] * us
            lists["blockages_pA"][i] = bl
            lists["stdev_pA"][i] = s[i, :k].copy()
        ts = t_start[ok]
        delay = np.empty(n)
        if n:
            delay[0] = ts[0]
            delay[1:] = ts[1:] - ts[:-1]                                   # mosaicConverter.py:73-76
        events = {"id": ids[ok], "type": np.zeros(n, np.int64), "start_time_s": ts, "event_delay_s": delay,
                  "duration_us": (ends[ok] - starts[ok]) * us, "threshold": np.full(n, float(threshold)),
                  "baseline_before_pA": before, "baseline_after_pA": after, "effective_baseline_pA": eff,
                  "area_pC": avg_block * dur_in / fs, "average_blockage_pA": avg_block,
                  "relative_average_blockage": avg_block / aeff, "max_blockage_pA": max_block,
                  "relative_max_blockage": max_block / aeff, "max_blockage_duration_us": length[rows, imax] * us,
                  "n_levels": L - 1, "intra_crossings": ic[ok], "rc_const1_us": np.zeros(n),
                  "rc_const2_us": np.zeros(n), "residual_pA": residual, "max_deviation_pA": max_dev,
                  "min_blockage_pA": min_block, "relative_min_blockage": min_block / aeff,
                  "min_blockage_duration_us": length[rows, imin] * us, **lists}
    return EventTable(events=events, rate=rate)


def event_table_from_result(an, r, *, samplerate: float, time_offset_s: float = 0.0, index_offset: int = 0) -> EventTable:
    """Event table of one `pipeline.TraceAnalyzer.run` result (device -> host, then
    `build_event_table`)."""
    lo, hi = event_extrema(r.detect_trace, r.win_start, r.win_end)
    tabs = an.tables_to_host(r)
    return build_event_table(starts=tabs["starts"], ends=tabs["ends"], types=tabs["types"], n_levels=tabs["n_levels"],
                             edges=tabs["edges"], level_mean=tabs["mean"], level_std=tabs["std"],
                             overflow=tabs["overflow"], xmin=lo.cpu().numpy(), xmax=hi.cpu().numpy(),
                             samplerate=samplerate, threshold=an.threshold, baseline_mean=r.baseline.mean,
                             baseline_std=r.baseline.std, baseline_block=an.block, padding=an.event_padding,
                             first_id=r.first_event_id, time_offset_s=time_offset_s, index_offset=index_offset,
                             block_offset=r.lo_halo, intra_count=tabs.get("intra_count"), intra_pairs=tabs.get("intra_pairs"))


def event_table_from_stream(san, rs, *, samplerate: float, time_offset_s: float = 0.0, index_offset: int = 0) -> EventTable:
    """Event table of one `pipeline.StreamingAnalyzer.run_from_host` result: the tables are already on
    the host; only the per-event extrema are taken from the device-resident filtered trace (windows =
    event +- event_padding, clipped to the rank's owned samples)."""
    t = rs.tables
    pad, n = int(san.kw.get("event_padding", 100)), int(rs.filtered.numel())
    dev = rs.filtered.device
    w0 = torch.from_numpy(np.clip(t["starts"] - pad, 0, n)).to(dev)
    w1 = torch.from_numpy(np.clip(t["ends"] + pad, 0, n)).to(dev)
    lo, hi = event_extrema(rs.filtered, w0, w1)
    return build_event_table(starts=t["starts"], ends=t["ends"], types=t["types"], n_levels=t["n_levels"],
                             edges=t["edges"], level_mean=t["mean"], level_std=t["std"], overflow=t["overflow"],
                             xmin=lo.cpu().numpy(), xmax=hi.cpu().numpy(), samplerate=samplerate,
                             threshold=float(san.kw.get("threshold", 5.0)), baseline_mean=rs.baseline.mean,
                             baseline_std=rs.baseline.std, baseline_block=san.block, padding=pad,
                             first_id=rs.first_event_id, time_offset_s=time_offset_s, index_offset=index_offset,
                             intra_count=t.get("intra_count"), intra_pairs=t.get("intra_pairs"))


def _write_csv(path: str, columns, table: dict) -> None:
    import pandas as pd
    df = pd.DataFrame({c: table[c] for c in columns}, columns=columns)
    for c in ("level_current_pA", "level_duration_us", "blockages_pA", "stdev_pA"):
        if c in df.columns and df[c].dtype == object:
            df[c] = [_fmt_list(v) for v in df[c]]
    df.to_csv(path, index=False, encoding="utf-8")


def write_analysis_dir(path: str, table: EventTable, *, baseline_mean, baseline_std, baseline_block: int,
                       samplerate: float, threshold: float, hysteresis: float, cutoff: float, poles: int,
                       extra_summary: dict | None = None, event_samples=None, time_offset_s: float = 0.0,
                       append: bool = False, intra_threshold: float = 0.0, intra_hysteresis: float = 0.0) -> None:
    """Write `events.csv`, `rate.csv`, `baseline.csv`, `summary.txt` and (if `event_samples`
    is given) `events/event_%08d.csv` under `path`.

    `event_samples` maps event id -> (window_start_index, samples float array, edges, level
    means): the rows of the per-event file are `time_us, current_pA, cusum_fit`
    (`readevents.py:1334-1337` names them time, current, cusum) WITH a header line, because
    the consumer reads them with an implicit header row (`readevents.py:1319`)."""
    os.makedirs(os.path.join(path, "events"), exist_ok=True)
    _write_csv(os.path.join(path, "events.csv"), EVENT_COLUMNS, table.events)
    _write_csv(os.path.join(path, "rate.csv"), RATE_COLUMNS, table.rate)
    fs = float(samplerate)
    nb = len(baseline_mean)
    _write_csv(os.path.join(path, "baseline.csv"), ["time_s", "baseline_pA", "stdev_pA"],
               {"time_s": time_offset_s + np.arange(nb) * (int(baseline_block) / fs),
                "baseline_pA": np.asarray(baseline_mean, np.float64), "stdev_pA": np.asarray(baseline_std, np.float64)})
    # summary.txt: consumers parse by SUBSTRING (plot-trace.py:190-203, readevents.py:73-79), so no
    # other key may contain 'threshold', 'hysteresis', 'cutoff' or 'poles'
    summary = {"threshold": repr(float(threshold)), "hysteresis": repr(float(hysteresis)), "cutoff": str(int(cutoff)),
               "poles": str(int(poles)), "intra_threshold": repr(float(intra_threshold)) if intra_threshold else "0",
               "intra_hysteresis": repr(float(intra_hysteresis)) if intra_hysteresis else "0", "samplerate": repr(fs),
               "baseline_block_samples": str(int(baseline_block)), "events_detected": str(len(table.rate["id"])),
               "events_accepted": str(len(table))}
    for k, v in (extra_summary or {}).items():
        if any(w in k for w in ("threshold", "hysteresis", "cutoff", "poles")):
            raise ValueError(f"summary key {k!r} would be mis-parsed by the consumers' substring match")
        summary[k] = str(v)
    # the two intra_ keys go last and the plain keys first: plot-trace.py skips lines containing 'intra'
    with open(os.path.join(path, "summary.txt"), "w") as f:
        for k, v in summary.items():
            f.write(f"{k}={v}\n")
    if event_samples:
        us = 1e6 / fs
        for eid, (w0, x, ed, mu) in event_samples.items():
            x = np.asarray(x, np.float64)
            fit = np.empty_like(x)
            for i in range(len(ed) - 1):
                fit[ed[i]:ed[i + 1]] = mu[i]
            t = np.arange(x.size) * us
            with open(os.path.join(path, "events", "event_%08d.csv" % int(eid)), "w") as f:
                f.write("time_us,current_pA,cusum_fit\n")
                np.savetxt(f, np.c_[t, x, fit], delimiter=",", fmt="%.16g")


def gather_event_samples(an, r, ids, table: EventTable) -> dict:
    """Samples + step fit of the chosen event ids (global ids) for the per-event files."""
    tabs = {k: v for k, v in zip(("w0", "w1"), (r.win_start.cpu().numpy(), r.win_end.cpu().numpy()))}
    nl = r.levels.n_levels.cpu().numpy(); ed = r.levels.edges.cpu().numpy(); mu = r.levels.mean.cpu().numpy()
    out = {}
    for eid in ids:
        i = int(eid) - r.first_event_id
        if i < 0 or i >= len(nl) or nl[i] <= 0:
            continue
        a, b = int(tabs["w0"][i]), int(tabs["w1"][i])
        out[int(eid)] = (a, r.detect_trace[a:b].cpu().numpy(), ed[i, :nl[i] + 1].astype(np.int64), mu[i, :nl[i]])
    return out
