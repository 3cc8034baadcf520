"""GPU parity of the loaders against the oracle restatement of the reference's loaders
(plot-trace.py:220-307, print_trace.py:32-39, legacy/minimal_psd.py:188-193) on files written
in the reference's formats, and of the float median behind np.pad(mode='median')."""
import os

import numpy as np
import pytest
import scipy.io as sio
import torch

from cusumtools_b200 import filters, loader, synth
from oracle import trace_oracle as to

pytestmark = pytest.mark.gpu


def write_series(tmp_path, gains, n_each=50_000, seed=5):
    rng = np.random.default_rng(seed)
    paths = []
    for i, g in enumerate(gains):
        name = tmp_path / f"run_201901{10 + i:02d}_1200{i:02d}.log"
        codes = (rng.integers(0, 16384, n_each).astype(np.uint16) << 2).astype(np.uint16)
        codes.tofile(name)
        st = dict(synth.CHIMERA_SETTINGS)
        st["SETUP_TIAgain"] = g
        sio.savemat(str(name).replace(".log", ".mat"), st)
        paths.append(str(name))
    return paths


def test_chimera_series_single_gain_codes(tmp_path):
    paths = write_series(tmp_path, [100e6, 100e6, 100e6])
    s = loader.ChimeraSeries(paths[1])
    assert [os.path.basename(p) for p in s.sorted_files] == [os.path.basename(p) for p in to.get_filenames(paths[1])]
    fs = s.samplerate
    for t0, t1 in ((0.0, 0.005), (0.010, 0.030), (0.0119, 0.0121), (0.02, 1.0)):
        raw, settings = s.load_codes(t0, t1)
        want = to.load_mapped_data(paths[0], t0, t1)[0]
        got = filters.scale_codes_host(raw.cpu().numpy(), settings)
        assert np.array_equal(got, want)


def test_chimera_series_mixed_gain_pA(tmp_path):
    paths = write_series(tmp_path, [100e6, 50e6, 100e6])
    s = loader.ChimeraSeries(paths[0])
    with pytest.raises(ValueError):
        s.load_codes(0.0, 0.03)
    got = s.load_pA(0.005, 0.030).cpu().numpy()
    want = to.load_mapped_data(paths[0], 0.005, 0.030)[0]
    assert got.shape == want.shape
    w32 = want.astype(np.float32)                            # float64 affine rounded once
    assert np.all(np.abs(got - w32) <= np.spacing(np.abs(w32)))
    assert np.mean(got == w32) > 0.9999


def test_bin_and_legacy_records(tmp_path):
    rng = np.random.default_rng(1)
    n = 30_001
    rec = np.zeros(n, dtype=np.dtype([("curr_pA", ">f8"), ("volt_mV", ">f8")]))
    rec["curr_pA"] = 5000 + 100 * rng.standard_normal(n)
    rec["volt_mV"] = 200.0
    p = tmp_path / "trace.bin"
    rec.tofile(p)
    got = loader.load_bin(str(p), 0.001, 0.004, 4166666.0).cpu().numpy()
    want = to.load_bin(str(p), 0.001, 0.004, 4166666.0)
    assert np.array_equal(got, want.astype(np.float32))
    assert np.array_equal(loader.load_bin(str(p)).cpu().numpy(), rec["curr_pA"].astype(np.float32))
    leg = np.zeros(n, dtype=np.dtype([("current", ">i2"), ("voltage", ">i2")]))
    leg["current"] = rng.integers(-32768, 32767, n)
    leg["voltage"] = 7
    q = tmp_path / "legacy.dat"
    leg.tofile(q)
    got = loader.load_legacy_i2(str(q), 100, 20_000, 0.25).cpu().numpy()
    want = to.load_legacy_i2(str(q), 100, 20_000, 0.25)
    assert np.array_equal(got, want.astype(np.float32))


@pytest.mark.parametrize("n", [1, 2, 1001, 1002, 300_000])
def test_float_median_exact(n):
    rng = np.random.default_rng(n)
    x = (5000 + 150 * rng.standard_normal(n)).astype(np.float32)
    x[rng.integers(0, n, max(1, n // 10))] *= -1
    t = torch.from_numpy(x).cuda()
    assert filters.float_median(t) == float(np.median(x.astype(np.float64)))
    assert filters.float_median(t, use_abs=True) == float(np.median(np.abs(x).astype(np.float64)))


def test_filter_data_on_bin_trace_takes_the_median_pad(tmp_path):
    rng = np.random.default_rng(2)
    x = (5000 + 150 * rng.standard_normal(200_000)).astype(np.float32)
    x[50_000:52_000] -= 800
    y = filters.bessel_filtfilt(torch.from_numpy(x).cuda(), 4166666.0, 1e5, 8).cpu().numpy()
    want = to.filter_data(x.astype(np.float64), 4166666.0, 1e5, 8)
    assert np.abs(y - want).max() < 0.05


def test_streaming_from_file_equals_streaming_from_host(tmp_path):
    """plot-trace.py:230-299 reads the window of a multi-file `.log` series before anything else happens; here the
    files are read piece by piece on a worker thread into rotating pinned slabs while earlier pieces are already
    being copied and analysed.  Same sub-shards, same kernels: the result equals the run from one host buffer."""
    import scipy.io as sio
    from cusumtools_b200 import pipeline
    codes, _ = synth.c1_trace(n=3_000_000, n_events=700, seed=21)
    cuts = [0, 1_100_000, 1_900_001, len(codes)]
    stamps = ["20210301_120000", "20210301_120001", "20210301_120002"]
    for i, st in enumerate(stamps):
        base = os.path.join(str(tmp_path), "pore_" + st)
        codes[cuts[i]:cuts[i + 1]].tofile(base + ".log")
        sio.savemat(base + ".mat", synth.CHIMERA_SETTINGS)
    series = loader.ChimeraSeries(os.path.join(str(tmp_path), "pore_" + stamps[1] + ".log"))
    fs = series.samplerate
    start_s, end_s = 100_000 / fs + 1e-9, 2_900_000 / fs + 1e-9
    rd = series.reader(start_s, end_s)
    lo, hi = int(start_s * fs), int(end_s * fs)
    assert rd.n == hi - lo
    kw = dict(threshold=5.0, hysteresis=1.0, baseline_block=65536, baseline_min=4700.0, baseline_max=5300.0,
              cusum_delta=400.0, cusum_h=10.0)
    an = pipeline.StreamingAnalyzer(rd.n, rd.settings, 1e5, 8, shards=5, **kw)
    from_file = an.run_from_file(rd, threads=3, slabs=2)
    tf = {k: v.copy() for k, v in from_file.tables.items()}
    yf = from_file.filtered.clone()
    from_host = an.run_from_host(torch.from_numpy(codes[lo:hi].copy()).pin_memory())
    assert from_file.median_codes == from_host.median_codes and from_file.total_events == from_host.total_events > 600
    valid = np.arange(tf["mean"].shape[1])[None, :] < tf["n_levels"][:, None]      # level rows are defined up to n_levels
    for k in tf:
        if k in ("mean", "std"):
            assert np.array_equal(tf[k][valid], from_host.tables[k][valid]), k
        else:
            assert np.array_equal(tf[k], from_host.tables[k]), k
    assert torch.equal(yf, from_host.filtered)
