"""Intra-event threshold crossings: the oracle definition against the literal state machine
on CPU, the CUDA kernel against the oracle on the GPU (bit-exact indices), and the columns the consumer
reads (readevents.py:1340-1343: parse_list of rate.csv's intra_crossing_times_us)."""
import numpy as np
import pytest

from oracle import events_oracle as eo


def literal(x, b, s, thr, hyst):
    """Walk the samples with the two lines the consumer draws (readevents.py:1363-1366)."""
    sign = 1.0 if b >= 0 else -1.0
    ts = np.float32(b - sign * thr * s)
    te = np.float32(b - sign * (thr - hyst) * s)
    out, inside, start = [], False, -1
    for i, v in enumerate(np.asarray(x, np.float32)):
        beyond = v < ts if sign > 0 else v > ts
        back = v > te if sign > 0 else v < te
        if not inside and beyond:
            inside, start = True, i
        elif inside and back:
            inside = False
            out.append((start, i))
    if inside:
        out.append((start, len(x)))
    return np.array(out, np.int64).reshape(-1, 2)


@pytest.mark.parametrize("b", [5000.0, -5000.0])
def test_oracle_equals_the_literal_state_machine(b):
    rng = np.random.default_rng(3)
    sign = 1.0 if b >= 0 else -1.0
    for trial in range(40):
        n = int(rng.integers(1, 700))
        x = b + rng.normal(0, 30, n)
        for _ in range(int(rng.integers(0, 4))):            # a few sub-events, some reaching the window end
            a = int(rng.integers(0, n)); w = int(rng.integers(1, 120))
            x[a:a + w] -= sign * rng.uniform(200, 900)
        got = eo.intra_crossings(x, b, 30.0, 8.0, 2.0)
        want = literal(x, b, 30.0, 8.0, 2.0)
        assert np.array_equal(got, want), trial
    assert eo.intra_crossings(np.zeros(0), b, 30.0, 8.0, 2.0).shape == (0, 2)


@pytest.mark.gpu
@pytest.mark.parametrize("negative", [False, True])
def test_kernel_equals_the_oracle(negative):
    import torch
    from cusumtools_b200 import detect
    rng = np.random.default_rng(11)
    n, block, E = 400_000, 65536, 600
    sgn = -1.0 if negative else 1.0
    y = (sgn * 5000.0 + rng.normal(0, 25, n)).astype(np.float32)
    starts = np.sort(rng.choice(np.arange(1000, n - 3000), E, replace=False)).astype(np.int64)
    lens = rng.integers(20, 2500, E)
    for s, l in zip(starts, lens):                           # events with deeper sub-events inside
        y[s:s + l] -= np.float32(sgn * 400.0)
        for _ in range(int(rng.integers(0, 6))):
            a = s + int(rng.integers(0, l)); w = int(rng.integers(1, 60))
            y[a:min(a + w, s + l + 50)] -= np.float32(sgn * rng.uniform(300, 800))
    w0 = np.maximum(starts - 100, 0); w1 = np.minimum(starts + lens + 100, n)
    yd = torch.from_numpy(y).cuda()
    lo, hi = (-5300.0, -4700.0) if negative else (4700.0, 5300.0)
    bl = detect.baseline_blocks(yd, block, lo, hi, threshold=5.0, hysteresis=1.0)
    K = 3
    nd = torch.tensor([E - 7], dtype=torch.int64, device="cuda")      # the last 7 rows must stay untouched
    cnt, pairs = detect.intra_crossings(yd, torch.from_numpy(w0).cuda(), torch.from_numpy(w1).cuda(),
                                        torch.from_numpy(starts).cuda(), bl, 20.0, 4.0, max_pairs=K, n_events_dev=nd)
    cnt, pairs = cnt.cpu().numpy(), pairs.cpu().numpy()
    mean, std = bl.mean, bl.std
    some = 0
    for i in range(E - 7):
        kb = min(starts[i] // block, len(mean) - 1)
        want = eo.intra_crossings(y[w0[i]:w1[i]], mean[kb], std[kb], 20.0, 4.0)
        assert cnt[i] == len(want), i
        k = min(len(want), K)
        assert np.array_equal(pairs[i, :2 * k].reshape(-1, 2), want[:k]), i
        some += len(want) > K
    assert cnt[:E - 7].sum() > E and some > 0                # the data exercise crossings and the truncation
    assert np.all(cnt[E - 7:] == 0) and np.all(pairs[E - 7:] == -1)
    # the detector's own lines were not disturbed
    assert bl.threshold == 5.0 and np.array_equal(bl.t_start, eo.thresholds(mean, std, 5.0, 1.0)[1])


@pytest.mark.gpu
def test_analyzer_and_writer_columns(tmp_path):
    import re
    import pandas as pd
    import torch
    from cusumtools_b200 import pipeline, synth, writer
    S = synth.CHIMERA_SETTINGS
    codes, _ = synth.c1_trace(n=1_500_000, n_events=300, seed=5)     # two-level events: the deeper level crosses
    kw = dict(threshold=5.0, hysteresis=1.0, baseline_min=4700.0, baseline_max=5300.0, baseline_block=65536,
              cusum_delta=400.0, cusum_h=10.0, intra_threshold=50.0, intra_hysteresis=5.0)
    an = pipeline.TraceAnalyzer(len(codes), S, 1e5, 8, **kw)
    r = an.run(torch.from_numpy(codes).cuda())
    t = writer.event_table_from_result(an, r, samplerate=synth.FS)
    cnt, pairs = r.intra[0].cpu().numpy(), r.intra[1].cpu().numpy()
    y = r.detect_trace.cpu().numpy()
    w0, w1 = r.win_start.cpu().numpy(), r.win_end.cpu().numpy()
    st = r.events.starts.cpu().numpy()
    for i in range(0, len(cnt), 7):
        kb = min(st[i] // 65536, len(r.baseline) - 1)
        want = eo.intra_crossings(y[w0[i]:w1[i]], r.baseline.mean[kb], r.baseline.std[kb], 50.0, 5.0)
        assert cnt[i] == len(want) and np.array_equal(pairs[i, :2 * len(want)].reshape(-1, 2), want)
    assert (cnt == 1).mean() > 0.9                                    # one deeper level per event
    assert np.array_equal(t.events["intra_crossings"], cnt[t.rate["type"] == 0])
    writer.write_analysis_dir(str(tmp_path), t, baseline_mean=r.baseline.mean, baseline_std=r.baseline.std,
                              baseline_block=65536, samplerate=synth.FS, threshold=5.0, hysteresis=1.0, cutoff=1e5, poles=8,
                              intra_threshold=50.0, intra_hysteresis=5.0)
    # the consumer's own statements: readevents.py:73-79 (summary keys), :1310 (parse_list), :1343 (pairs)
    it = ih = 0.0
    for line in open(tmp_path / "summary.txt"):
        if "intra_threshold" in line:
            it = float(re.split("=|\n", line)[1])
        if "intra_hysteresis" in line:
            ih = float(re.split("=|\n", line)[1])
    assert it == 50.0 and ih == 5.0
    ratedb = pd.read_csv(tmp_path / "rate.csv", encoding="utf-8")
    i = int(np.nonzero(cnt == 1)[0][0])
    semilist = np.squeeze(ratedb.loc[ratedb["id"] == i, "intra_crossing_times_us"].values)
    crossings = np.hstack([np.array(a, dtype=float) for a in str(semilist).split(";")]).astype(np.float64)
    got = list(zip(crossings[::2], crossings[1::2]))
    assert len(got) == 1 and np.allclose(got[0], pairs[i, :2] * 1e6 / synth.FS)
    # streamed form carries the same columns
    sa = pipeline.StreamingAnalyzer(len(codes), S, 1e5, 8, shards=3, **kw)
    rs = sa.run_from_host(torch.from_numpy(codes).pin_memory())
    ts = writer.event_table_from_stream(sa, rs, samplerate=synth.FS)
    assert np.array_equal(ts.events["intra_crossings"], t.events["intra_crossings"])
