"""The analysis directory (events.csv / rate.csv / baseline.csv / summary.txt /
events/event_%08d.csv) round-trips through the reference consumers' own parsing statements
(readevents.py:73-79,843-854,1297-1306,1318-1337; plot-trace.py:181-203,354-357,385-398),
and every column follows the in-repo field definitions (mosaicConverter.py:72-154) when
recomputed directly from the samples.  CPU only: the level tables come from the oracle."""
import os
import re

import numpy as np
import pandas as pd

from cusumtools_b200 import synth, writer
from oracle import c_twin, events_oracle as eo

FS = 4166666.0
BLOCK, PAD = 65536, 100


def make_tables(n=400_000, n_events=90, seed=4, ML=16):
    rng = np.random.default_rng(seed)
    y = (5000 + 24 * rng.standard_normal(n)).astype(np.float32)
    for k in range(n_events):
        s = 2000 + 4000 * k
        y[s:s + 1000] -= 800
        y[s + 1000:s + 2000] -= 1600
    y[5000:5004] -= 900                                # too short: rejected (type 2)
    c0 = np.float32(5000.0); sh = eo.stats_shift(300.0, BLOCK)
    mean, std = eo.baseline_from_stats(*c_twin.block_stats(y, BLOCK, 4700.0, 5300.0, c0, sh), c0, sh)
    s, e, _ = c_twin.detect_events(y, BLOCK, *eo.thresholds(mean, std, 5.0, 1.0))
    w0, w1, typ = eo.event_windows(s, e, n, PAD, 8, 100000)
    ok = typ == 0
    offs = np.concatenate(([0], np.cumsum((w1 - w0)[ok])))
    flat = np.concatenate([y[a:b] for a, b in zip(w0[ok], w1[ok])])
    nl_ok, ed_ok, mu_ok, sd_ok, ov_ok = c_twin.cusum_batch(flat, offs, 400.0, 10.0, ML)
    E = len(s)
    nl = np.zeros(E, np.int32); ed = np.full((E, ML + 1), -1, np.int32); mu = np.zeros((E, ML)); sd = np.zeros((E, ML))
    ov = np.zeros(E, np.uint8)
    nl[ok], ed[ok], mu[ok], sd[ok], ov[ok] = nl_ok, ed_ok, mu_ok, sd_ok, ov_ok
    xmin = np.array([y[max(a, 0):b].min() for a, b in zip(w0, w1)], np.float32)
    xmax = np.array([y[max(a, 0):b].max() for a, b in zip(w0, w1)], np.float32)
    tab = writer.build_event_table(starts=s, ends=e, types=typ, n_levels=nl, edges=ed, level_mean=mu, level_std=sd,
                                   overflow=ov, xmin=xmin, xmax=xmax, samplerate=FS, threshold=5.0,
                                   baseline_mean=mean, baseline_std=std, baseline_block=BLOCK, padding=PAD)
    return y, tab, dict(s=s, e=e, w0=w0, w1=w1, typ=typ, nl=nl, ed=ed, mu=mu, mean=mean, std=std)


def test_columns_follow_the_field_definitions():
    y, tab, d = make_tables()
    ev = tab.events
    assert len(tab) >= 85 and list(ev.keys())[:3] == ["id", "type", "start_time_s"]
    assert set(writer.EVENT_COLUMNS) == set(ev.keys())
    assert np.all(np.diff(ev["id"]) > 0)
    assert 2 in tab.rate["type"]                       # the 4-sample glitch is in rate.csv only
    us = 1e6 / FS
    for row in range(0, len(tab), 7):
        i = int(ev["id"][row])
        w0, nl, ed = d["w0"][i], d["nl"][i], d["ed"][i]
        x = y[w0:d["w1"][i]].astype(np.float64)
        # levels recomputed from the samples with numpy (mosaicConverter.py:113-130 uses np.mean / np.std)
        lv = [x[ed[k]:ed[k + 1]] for k in range(nl)]
        cur = np.array([v.mean() for v in lv]); sdv = np.array([v.std() for v in lv])
        assert np.allclose(ev["level_current_pA"][row], cur, rtol=0, atol=2e-2)       # 1/64 pA quantisation
        assert np.allclose(ev["stdev_pA"][row], sdv, rtol=2e-3, atol=1e-2)
        assert np.allclose(ev["level_duration_us"][row], [len(v) * us for v in lv])
        eff = 0.5 * (cur[0] + cur[-1])
        assert abs(ev["effective_baseline_pA"][row] - eff) < 2e-2
        inner = x[ed[1]:ed[nl - 1]]
        assert abs(ev["average_blockage_pA"][row] - (eff - inner.mean())) < 3e-2     # mosaicConverter.py:113
        assert abs(ev["area_pC"][row] - (eff - inner).sum() / FS) < 1e-6             # pA * s
        assert abs(ev["max_deviation_pA"][row] - np.abs(x - ev["effective_baseline_pA"][row]).max()) < 1e-3
        fit = np.concatenate([np.full(len(v), c) for v, c in zip(lv, cur)])
        assert abs(ev["residual_pA"][row] - np.std(x - fit)) < 2e-2                  # mosaicConverter.py:104
        b = ev["blockages_pA"][row]
        assert abs(b[0] - (cur[0] - eff)) < 2e-2 and abs(b[-1] - (cur[-1] - eff)) < 2e-2
        assert ev["n_levels"][row] == nl - 1 and nl >= 3
        assert abs(ev["max_blockage_pA"][row] - max(b[1:-1])) < 1e-9 and abs(ev["min_blockage_pA"][row] - min(b[1:-1])) < 1e-9
        assert abs(ev["duration_us"][row] - (d["e"][i] - d["s"][i]) * us) < 1e-9
    assert abs(ev["event_delay_s"][0] - ev["start_time_s"][0]) < 1e-15
    assert np.allclose(ev["event_delay_s"][1:], np.diff(ev["start_time_s"]))
    assert np.allclose(ev["relative_max_blockage"], ev["max_blockage_pA"] / np.abs(ev["effective_baseline_pA"]))
    # the injected template: two sub-levels of 800 and 1600 pA
    assert np.median(ev["n_levels"]) == 3 and abs(np.median(ev["max_blockage_pA"]) - 1600) < 10


def test_directory_round_trips_through_the_consumers(tmp_path):
    y, tab, d = make_tables(n=200_000, n_events=40)
    out = str(tmp_path / "analysis")
    first = int(tab.events["id"][0])
    i = first
    samples = {first: (int(d["w0"][i]), y[d["w0"][i]:d["w1"][i]], d["ed"][i, :d["nl"][i] + 1].astype(np.int64),
                       d["mu"][i, :d["nl"][i]])}
    writer.write_analysis_dir(out, tab, baseline_mean=d["mean"], baseline_std=d["std"], baseline_block=BLOCK,
                              samplerate=FS, threshold=5.0, hysteresis=1.0, cutoff=100000.0, poles=8,
                              extra_summary={"cusum_delta": 400.0, "cusum_h": 10.0}, event_samples=samples)
    # ---- readevents.main (readevents.py:1527-1544)
    eventsdb = pd.read_csv(os.path.join(out, "events.csv"), encoding="utf-8")
    ratedb = pd.read_csv(os.path.join(out, "rate.csv"), encoding="utf-8")
    assert list(eventsdb.columns) == writer.EVENT_COLUMNS and len(eventsdb) == len(tab)
    # readevents.py:73-79
    intra_threshold = intra_hysteresis = None
    for line in open(os.path.join(out, "summary.txt")):
        if "intra_threshold" in line:
            intra_threshold = float(re.split("=|\n", line)[1])
        if "intra_hysteresis" in line:
            intra_hysteresis = float(re.split("=|\n", line)[1])
    assert intra_threshold == 0 and intra_hysteresis == 0
    # readevents.py:843-846 first_level_fraction, :850-851 folding
    durations = [np.array(a, dtype=float)[1:-1] for a in eventsdb["level_duration_us"].str.split(";")]
    fraction = [du[0] / (np.sum(du) + du[0]) for du in durations]
    assert all(0 < f < 1 for f in fraction)
    folding = eventsdb["max_blockage_duration_us"] / (eventsdb["duration_us"] + eventsdb["max_blockage_duration_us"])
    assert np.all((folding > 0) & (folding < 1))
    # readevents.py:1297-1306 parse_db_col with and without the baseline entries
    for colname in ("blockages_pA", "level_current_pA", "level_duration_us", "stdev_pA"):
        full = np.hstack([np.array(a, dtype=float) for a in eventsdb[colname].str.split(";")])
        core = np.hstack([np.array(a, dtype=float)[1:-1] for a in eventsdb[colname].str.split(";")])
        assert full.size == core.size + 2 * len(eventsdb) and np.all(np.isfinite(full))
    assert np.allclose(np.hstack([np.array(a, dtype=float) for a in eventsdb["level_current_pA"].str.split(";")]),
                       np.hstack(list(tab.events["level_current_pA"])), rtol=1e-15, atol=0)   # '%.16g' as the reference
    # readevents.py:1318-1337: per-event file read with an implicit header, type 0 -> 3 columns
    event_file = pd.read_csv(os.path.join(out, "events", "event_%08d.csv" % first), encoding="utf-8")
    event_file.columns = ["time", "current", "cusum"]
    assert len(event_file) == d["w1"][i] - d["w0"][i]                                # no sample swallowed
    assert np.allclose(event_file["current"].values, y[d["w0"][i]:d["w1"][i]])
    assert abs(event_file["time"].values[1] - 1e6 / FS) < 1e-9
    # readevents.py:1340-1343 columns of rate.csv
    for c in ("intra_crossing_times_us", "local_stdev", "local_baseline"):
        assert c in ratedb.columns
    # ---- plot-trace.overlay_cusum (plot-trace.py:181-203)
    threshold = hysteresis = config_cutoff = config_order = None
    with open(os.path.join(out, "summary.txt"), "r") as config:
        for line in config:
            if "threshold" in line and "intra" not in line:
                threshold = float(re.split("=|\n", line)[1])
            if "hysteresis" in line and "intra" not in line:
                hysteresis = float(re.split("=|\n", line)[1])
            if "cutoff" in line:
                config_cutoff = int(re.split("=|\n", line)[1])
            if "poles" in line:
                config_order = int(re.split("=|\n", line)[1])
    assert (threshold, hysteresis, config_cutoff, config_order) == (5.0, 1.0, 100000, 8)
    # plot-trace.py:354-357 (sqldf restated with pandas): good / rejected events in a time window
    good = ratedb[(ratedb.start_time_s >= 0) & (ratedb.start_time_s < 1) & ratedb.type.isin([0, 1])]
    bad = ratedb[(ratedb.start_time_s >= 0) & (ratedb.start_time_s < 1) & (ratedb.type > 1)]
    assert len(good) == len(tab) and len(good) + len(bad) == len(ratedb)
    assert np.all(good.end_time_s.values > good.start_time_s.values)
    # plot-trace.py:385-398
    base = pd.read_csv(os.path.join(out, "baseline.csv"), encoding="utf-8")
    assert list(base.columns) == ["time_s", "baseline_pA", "stdev_pA"] and len(base) == len(d["mean"])
    assert np.allclose(base["time_s"].values, np.arange(len(base)) * BLOCK / FS)
    assert np.array_equal(base["baseline_pA"].values, d["mean"])


def test_summary_rejects_keys_the_consumers_would_misparse(tmp_path):
    _, tab, d = make_tables(n=100_000, n_events=10)
    import pytest
    with pytest.raises(ValueError):
        writer.write_analysis_dir(str(tmp_path / "x"), tab, baseline_mean=d["mean"], baseline_std=d["std"],
                                  baseline_block=BLOCK, samplerate=FS, threshold=5.0, hysteresis=1.0, cutoff=1e5,
                                  poles=8, extra_summary={"cusum_threshold": 10})


def test_intra_crossing_columns_round_trip_through_the_consumer(tmp_path):
    """rate.csv `intra_crossing_times_us` / events.csv `intra_crossings` / summary keys from the oracle's
    crossings, read back with the consumer's statements (readevents.py:73-79, 1310, 1340-1343)."""
    y, _, d = make_tables()
    E = len(d["s"])
    K = 4
    cnt = np.zeros(E, np.int32); pairs = np.full((E, 2 * K), -1, np.int32)
    for i in range(E):
        kb = min(d["s"][i] // BLOCK, len(d["mean"]) - 1)
        c = eo.intra_crossings(y[max(d["w0"][i], 0):d["w1"][i]], d["mean"][kb], d["std"][kb], 50.0, 5.0)
        cnt[i] = len(c)
        pairs[i, :2 * min(len(c), K)] = c[:K].ravel()
    assert (cnt == 1).mean() > 0.9                        # the deeper second level of every synthetic event
    xmin = np.array([y[max(a, 0):b].min() for a, b in zip(d["w0"], d["w1"])], np.float32)
    xmax = np.array([y[max(a, 0):b].max() for a, b in zip(d["w0"], d["w1"])], np.float32)
    ov = np.zeros(E, np.uint8)
    sdv = np.zeros_like(d["mu"])
    tab = writer.build_event_table(starts=d["s"], ends=d["e"], types=d["typ"], n_levels=d["nl"], edges=d["ed"], level_mean=d["mu"],
                                   level_std=sdv, overflow=ov, xmin=xmin, xmax=xmax, samplerate=FS, threshold=5.0,
                                   baseline_mean=d["mean"], baseline_std=d["std"], baseline_block=BLOCK, padding=PAD,
                                   intra_count=cnt, intra_pairs=pairs)
    assert np.array_equal(tab.events["intra_crossings"], cnt[tab.rate["type"] == 0])
    writer.write_analysis_dir(str(tmp_path), tab, baseline_mean=d["mean"], baseline_std=d["std"], baseline_block=BLOCK,
                              samplerate=FS, threshold=5.0, hysteresis=1.0, cutoff=1e5, poles=8, intra_threshold=50.0,
                              intra_hysteresis=5.0)
    it = ih = thr = None
    for line in open(tmp_path / "summary.txt"):
        if "intra_threshold" in line:
            it = float(re.split("=|\n", line)[1])
        if "intra_hysteresis" in line:
            ih = float(re.split("=|\n", line)[1])
        if "threshold" in line and "intra" not in line:
            thr = float(re.split("=|\n", line)[1])
    assert (it, ih, thr) == (50.0, 5.0, 5.0)
    ratedb = pd.read_csv(tmp_path / "rate.csv", encoding="utf-8")
    for i in np.nonzero(cnt > 0)[0][:10]:
        semilist = np.squeeze(ratedb.loc[ratedb["id"] == i, "intra_crossing_times_us"].values)
        crossings = np.hstack([np.array(a, dtype=float) for a in str(semilist).split(";")]).astype(np.float64)
        got = np.array(list(zip(crossings[::2], crossings[1::2])))
        want = pairs[i, :2 * min(cnt[i], K)].reshape(-1, 2) * 1e6 / FS
        assert np.allclose(got, want, rtol=1e-12)
    evdb = pd.read_csv(tmp_path / "events.csv", encoding="utf-8")
    assert np.array_equal(evdb["intra_crossings"].values, tab.events["intra_crossings"])
