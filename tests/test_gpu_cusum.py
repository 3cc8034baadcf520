"""GPU parity: batched CUSUM+ kernel vs the oracle definition (NumPy / its plain-C twin)
on identical float32 samples: n_levels, changepoint edges, overflow flags AND the level
means / standard deviations are bit-exact (all reductions are over exact integers)."""
import numpy as np
import pytest
import torch

from cusumtools_b200 import cusum, synth
from oracle import c_twin, events_oracle as eo

pytestmark = pytest.mark.gpu


def run_gpu(x, offsets, delta, h, max_levels=16, types=None):
    t = cusum.cusum_flat(torch.from_numpy(x).cuda(), torch.from_numpy(offsets).cuda(), delta=delta, h=h,
                         max_levels=max_levels, types=types)
    torch.cuda.synchronize()
    return [v.cpu().numpy() for v in (t.n_levels, t.edges, t.mean, t.std, t.overflow)]


def assert_same(got, want):
    for name, a, b in zip(("n_levels", "edges", "mean", "std", "overflow"), got, want):
        assert np.array_equal(a, b), name


def test_c3_sample_bit_exact():
    x, offsets, nlev = synth.c3_events(3000, seed=2024)
    got = run_gpu(x, offsets, 400.0, 10.0)
    want = c_twin.cusum_batch(x, offsets, 400.0, 10.0, 16)
    assert_same(got, want)
    assert np.mean(got[0] == nlev + 2) > 0.9
    # the numpy definition itself on a subset
    sub = slice(0, 40)
    w = eo.cusum_batch(x[:offsets[40]], offsets[:41], 400.0, 10.0, 16)
    for a, b in zip(got, w):
        assert np.array_equal(a[sub], b)


@pytest.mark.parametrize("delta,h", [(200.0, 5.0), (400.0, 10.0), (800.0, 40.0), (50.0, 2.0)])
def test_parameter_sweep(delta, h):
    x, offsets, _ = synth.c3_events(400, seed=int(delta), max_len=3000)
    assert_same(run_gpu(x, offsets, delta, h), c_twin.cusum_batch(x, offsets, delta, h, 16))


def test_edge_cases():
    rng = np.random.default_rng(3)
    parts = [np.full(50, 7.0, np.float32),                       # zero variance: no information
             np.array([1.0], np.float32),                         # single sample
             np.array([1.0, 900.0], np.float32),                  # two samples
             (24 * rng.standard_normal(255)).astype(np.float32),  # one block minus one
             (24 * rng.standard_normal(256)).astype(np.float32),
             (24 * rng.standard_normal(257)).astype(np.float32)]
    y = (24 * rng.standard_normal(4000)).astype(np.float32)       # overflow: 19 jumps, max_levels 8
    for k in range(1, 20):
        y[k * 200:] += np.float32(1000 * (-1) ** k)
    parts.append(y)
    parts.append(np.zeros(0, np.float32))                         # empty window
    big = (5000 + 24 * rng.standard_normal(70000)).astype(np.float32)   # long event, many blocks
    big[30000:50000] -= 900
    parts.append(big)
    offsets = np.concatenate(([0], np.cumsum([len(p) for p in parts]))).astype(np.int64)
    x = np.concatenate(parts)
    got = run_gpu(x, offsets, 400.0, 10.0, max_levels=8)
    want = c_twin.cusum_batch(x, offsets, 400.0, 10.0, 8)
    assert_same(got, want)
    assert got[4][6] == 1 and got[0][7] == 0 and got[0][8] == 3


def test_unaligned_windows_inside_a_trace():
    """Windows into a long trace at arbitrary offsets (the detection -> CUSUM path)."""
    rng = np.random.default_rng(8)
    y = (5000 + 24 * rng.standard_normal(300000)).astype(np.float32)
    w0, w1 = [], []
    for k in range(60):
        s = 1000 + 4777 * k + int(rng.integers(0, 9))
        L = int(rng.integers(300, 2500))
        y[s:s + L // 2] -= 800
        y[s + L // 2:s + L] -= 1500
        w0.append(s - 100); w1.append(s + L + 100)
    w0 = np.array(w0, np.int64); w1 = np.array(w1, np.int64)
    typ = np.zeros(60, np.int32); typ[5] = 3
    t = cusum.cusum_levels(torch.from_numpy(y).cuda(), torch.from_numpy(w0).cuda(), torch.from_numpy(w1).cuda(),
                           delta=400.0, h=10.0, types=torch.from_numpy(typ).cuda())
    torch.cuda.synchronize()
    flat = np.concatenate([y[a:b] for a, b in zip(w0, w1)])
    offs = np.concatenate(([0], np.cumsum(w1 - w0)))
    want = c_twin.cusum_batch(flat, offs, 400.0, 10.0, 16)
    nl = t.n_levels.cpu().numpy()
    assert nl[5] == 0
    keep = typ == 0
    assert np.array_equal(nl[keep], want[0][keep])
    assert np.array_equal(t.edges.cpu().numpy()[keep], want[1][keep])
    assert np.array_equal(t.mean.cpu().numpy()[keep], want[2][keep])
    assert np.array_equal(t.std.cpu().numpy()[keep], want[3][keep])
    assert np.all(nl[keep] == 4)


def test_full_c3_checksum_properties():
    """1M-event config at reduced event count on the device: order invariance (shuffling
    the event order permutes the outputs) and agreement with the C twin on a slice."""
    x, offsets, _ = synth.c3_events(100000, seed=99)
    xt = torch.from_numpy(x).cuda()
    w0 = torch.from_numpy(offsets[:-1].copy()).cuda(); w1 = torch.from_numpy(offsets[1:].copy()).cuda()
    a = cusum.cusum_levels(xt, w0, w1, delta=400.0, h=10.0)
    perm = torch.randperm(w0.numel(), device="cuda")
    b = cusum.cusum_levels(xt, w0[perm].contiguous(), w1[perm].contiguous(), delta=400.0, h=10.0)
    assert torch.equal(a.edges[perm], b.edges) and torch.equal(a.mean[perm], b.mean)
    want = c_twin.cusum_batch(x[:offsets[5000]], offsets[:5001], 400.0, 10.0, 16)
    assert np.array_equal(a.edges[:5000].cpu().numpy(), want[1])
    assert np.array_equal(a.std[:5000].cpu().numpy(), want[3])


def test_event_columns_kernel_matches_the_numpy_statement():
    """The non-level columns of events.csv (mosaicConverter.py:72-154: baselines, blockages, area inputs, residual,
    max deviation, final type) computed one thread per event on the device against writer.event_columns_host."""
    from cusumtools_b200 import pipeline, writer
    codes, _ = synth.c1_trace(n=1_500_000, n_events=350, seed=12)
    S = synth.CHIMERA_SETTINGS
    for ml in (16, 3):                        # 3: every two-level event overflows the table (type 6)
        an = pipeline.TraceAnalyzer(len(codes), S, 1e5, 8, threshold=5.0, hysteresis=1.0, baseline_block=65536,
                                    baseline_min=4700.0, baseline_max=5300.0, cusum_delta=400.0, cusum_h=10.0, max_levels=ml)
        r = an.run(torch.from_numpy(codes).cuda())
        lo, hi = writer.event_extrema(r.detect_trace, r.win_start, r.win_end)
        cols, typ = writer.event_columns(r.levels, r.types, lo, hi)
        want_c, want_t = writer.event_columns_host(r.types.cpu().numpy(), r.levels.n_levels.cpu().numpy(), r.levels.edges.cpu().numpy(),
                                                   r.levels.mean.cpu().numpy(), r.levels.std.cpu().numpy(),
                                                   r.levels.overflow.cpu().numpy(), lo.cpu().numpy(), hi.cpu().numpy())
        assert np.array_equal(typ.cpu().numpy(), want_t)
        assert (want_t == 0).sum() > 300 if ml == 16 else (want_t == 6).sum() > 300
        assert np.allclose(cols.cpu().numpy(), want_c, rtol=1e-12, atol=1e-9)
        tab = writer.event_table_from_result(an, r, samplerate=synth.FS)
        assert len(tab) == int((want_t == 0).sum())
        if ml == 16:
            assert abs(np.median(tab.events["max_blockage_pA"]) - 1600) < 100


def _windows_case(seed, n_events, ntot_pad, max_levels=8, unaligned=False):
    """>= 16384 windows (so that the thread-per-event kernel takes them) with every shape of window it has special code
    for: lengths 1, 2, 31, 32, 33 and other short ones, plateaus longer than the shared-memory reciprocal table (8192),
    windows longer than a lane takes (16384: left to the warps), a window ending on the last sample of a trace whose
    length is not a multiple of the 128-byte line, a window starting at sample 0, rejected types, level overflow."""
    rng = np.random.default_rng(seed)
    lens = rng.integers(40, 400, n_events)
    lens[:12] = [1, 2, 31, 32, 33, 63, 64, 65, 9000, 12000, 17000, 20000]
    rng.shuffle(lens)
    gaps = rng.integers(0, 70, n_events)
    w0 = np.cumsum(lens + gaps) - lens - gaps[0]            # first window starts at sample 0
    w1 = w0 + lens
    ntot = int(w1[-1]) + ntot_pad
    y = (5000 + 24 * rng.standard_normal(ntot)).astype(np.float32)
    for e in range(n_events):
        L = int(lens[e]); a = int(w0[e])
        if L >= 40:
            k = int(rng.integers(0, 4))
            cuts = np.linspace(a + 5, a + L - 5, k + 2).astype(np.int64)
            for i in range(k + 1):
                if i % 2 == 0 and k > 0:
                    y[cuts[i]:cuts[i + 1]] -= np.float32(600 + 400 * (i % 3))
        if e % 97 == 0 and L > 200:                           # many jumps: more levels than the table holds
            for j in range(12):
                y[a + 10 + 15 * j:a + L] += np.float32(900 * (-1) ** j)
    typ = np.zeros(n_events, np.int32); typ[rng.integers(0, n_events, 50)] = 2
    return y, w0.astype(np.int64), w1.astype(np.int64), typ, max_levels


@pytest.mark.parametrize("ntot_pad,unaligned", [(0, False), (13, False), (5, True)])
def test_thread_per_event_kernel_window_shapes(ntot_pad, unaligned):
    y, w0, w1, typ, ml = _windows_case(31 + ntot_pad, 20000, ntot_pad)
    buf = torch.empty(y.size + 8, dtype=torch.float32, device="cuda")
    yt = buf[1:1 + y.size] if unaligned else buf[:y.size]     # unaligned: no 16-byte copies, every window goes to the warps
    yt.copy_(torch.from_numpy(y))
    t = cusum.cusum_levels(yt, torch.from_numpy(w0).cuda(), torch.from_numpy(w1).cuda(), delta=400.0, h=10.0,
                           max_levels=ml, types=torch.from_numpy(typ).cuda())
    torch.cuda.synchronize()
    flat = np.concatenate([y[a:b] for a, b in zip(w0, w1)])
    offs = np.concatenate(([0], np.cumsum(w1 - w0)))
    want = c_twin.cusum_batch(flat, offs, 400.0, 10.0, ml)
    keep = typ == 0
    nl = t.n_levels.cpu().numpy()
    assert np.all(nl[~keep] == 0)
    assert np.array_equal(nl[keep], want[0][keep])
    assert np.array_equal(t.edges.cpu().numpy()[keep], want[1][keep])
    assert np.array_equal(t.mean.cpu().numpy()[keep], want[2][keep])
    assert np.array_equal(t.std.cpu().numpy()[keep], want[3][keep])
    assert np.array_equal(t.overflow.cpu().numpy()[keep], want[4][keep])
    assert want[4][keep].sum() > 5 and (want[0][keep] >= 3).sum() > 5000


@pytest.mark.parametrize("delta,h,noise", [(400.0, 10.0, 33.0), (50.0, 2.0, 24.0), (800.0, 40.0, 60.0), (120.0, 1.0, 5.0)])
def test_thread_per_event_kernel_on_ramps(delta, h, noise):
    """17 000 windows (the thread-per-event kernel) whose level changes are RAMPS of 0-20 samples, as on a filtered trace:
    the statistics stay non-zero for many samples, several changepoints follow each other, and with a small delta nearly
    every sample takes the full evaluation.  Bit-exact against the C twin."""
    rng = np.random.default_rng(int(delta + h))
    n_events = 17000
    lens = rng.integers(60, 500, n_events)
    offs = np.concatenate(([0], np.cumsum(lens))).astype(np.int64)
    x = 5000 + noise * rng.standard_normal(int(offs[-1]))
    for e in range(n_events):
        a, n = int(offs[e]), int(lens[e])
        for c in np.sort(rng.integers(10, n - 10, int(rng.integers(0, 4)))):
            depth = float(rng.choice([-1600.0, -800.0, 600.0, 1200.0]))
            ramp = int(rng.integers(0, 20))
            prof = np.ones(n - c) if ramp == 0 else np.minimum(1.0, (np.arange(n - c) + 1) / ramp)
            x[a + c:a + n] += depth * prof
    x = x.astype(np.float32)
    got = run_gpu(x, offs, delta, h, max_levels=12)
    want = c_twin.cusum_batch(x, offs, delta, h, 12)
    assert_same(got, want)
    assert (want[0] >= 3).sum() > 3000
