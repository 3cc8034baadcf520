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
