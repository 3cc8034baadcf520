"""GPU parity of the headless entry-point mirrors (cusumtools_b200/app.py) against the oracle restatement of the
same reference methods: the tests drive the classes the way the reference's GUI does (fill the entries, press
'Update Trace' / 'Update PSD') and compare what the reference would have plotted or exported.

Tolerances as everywhere: filtered samples 0.05 pA, PSD bins 5e-6 P + 1e-9 max P (one arbitrary-length
segment: 5e-5 / 1e-8), cumulative rms rel 1e-6 (1e-5), dequantised samples one float32 rounding."""
import os

import numpy as np
import pytest
import scipy.io as sio
import torch

from cusumtools_b200 import app as ctapp
from cusumtools_b200 import synth
from oracle import trace_oracle as to

pytestmark = pytest.mark.gpu

FILTER_TOL_PA = 0.05


def write_series(tmp_path, gains=(100e6, 100e6, 100e6), n_each=700_000, seed=3):
    """Three `.log` files of a noisy 5 nA trace with blockades, each with its own TIA gain."""
    paths = []
    for i, g in enumerate(gains):
        st = dict(synth.CHIMERA_SETTINGS)
        st["SETUP_TIAgain"] = g
        codes, _ = synth.c1_trace(n=n_each, n_events=n_each // 4000 - 2, seed=seed + i, settings=st)
        name = tmp_path / f"pore_20210305_1015{i:02d}.log"
        codes.tofile(name)
        sio.savemat(str(name).replace(".log", ".mat"), st)
        paths.append(str(name))
    return paths


def chain_input(a, want_filtered):
    """Stage-by-stage parity: the filter stage is held to its own tolerance against the reference's float64 call
    sequence, and the PSD stage is then compared on identical input (the float32 samples the GPU filter produced)."""
    y = a.filtered_data.cpu().numpy()
    assert y.shape == want_filtered.shape and np.abs(y - want_filtered).max() <= FILTER_TOL_PA
    return y.astype(np.float64)


def psd_close(P, want, single=False):
    tol = (5e-5 * want + 1e-8 * want.max()) if single else (5e-6 * want + 1e-9 * want.max())
    return not (np.abs(P - want) > tol).any()


@pytest.mark.parametrize("gains", [(100e6, 100e6, 100e6), (100e6, 50e6, 100e6)])
def test_update_trace_filtered_window_across_files(tmp_path, gains):
    paths = write_series(tmp_path, gains)
    a = ctapp.App(None, paths[1])
    a.start_entry.set("0.1")
    a.end_entry.set("0.45")                       # 0.168 s per file: first, middle and last file pieces
    a.cutoff_entry.set("100000")
    a.update_trace()
    data, fs = to.load_mapped_data(paths[0], 0.1, 0.45)
    got = a.data.cpu().numpy()
    assert got.shape == data.shape and np.all(np.abs(got - data.astype(np.float32)) <= np.spacing(np.abs(data.astype(np.float32))))
    want = to.filter_data(data, fs, 100000.0, 8)
    y = a.filtered_data.cpu().numpy()
    assert y.shape == want.shape and np.abs(y - want).max() <= FILTER_TOL_PA
    assert a.plot_data is a.filtered_data and a.plot_samplerate == fs
    t = a.plot_time_us()
    assert len(t) == len(want) and np.isclose(t[0], (1.0 / fs + 0.1) * 1e6) and np.isclose(t[-1], (len(want) / fs + 0.1) * 1e6)


def test_update_trace_unfiltered_and_downsampled(tmp_path):
    paths = write_series(tmp_path)
    a = ctapp.App(None, paths[0])
    a.start_entry.set("0")
    a.end_entry.set("0.2")
    a.cutoff_entry.set("")
    a.update_trace()
    assert a.filtered_data is a.data and a.plot_data is a.data         # plot-trace.py:332-333
    a.cutoff_entry.set("100000")
    a.order_entry.set("4")
    a.downsample_entry.set("500000")
    a.update_trace()
    data, fs = to.load_mapped_data(paths[0], 0.0, 0.2)
    want = to.filter_data(data, fs, 100000.0, 4)[::int(fs / 500000.0)]
    got = a.plot_data.cpu().numpy()
    assert got.shape == want.shape and np.abs(got - want).max() <= FILTER_TOL_PA and a.plot_samplerate == 500000.0
    a.export_trace(str(tmp_path / "trace.csv"))
    assert np.allclose(np.loadtxt(tmp_path / "trace.csv", delimiter=","), got.astype(np.float64), rtol=0, atol=0)


def test_update_trace_on_replaced_data_uses_the_float_path(tmp_path):
    """Code that assigns `app.data` itself (as the reference's attribute allows) is filtered from that data."""
    paths = write_series(tmp_path)
    a = ctapp.App(None, paths[0])
    a.end_entry.set("0.1")
    a.load_mapped_data()
    x = 3000.0 + 50.0 * np.random.default_rng(0).standard_normal(200_001)
    a.data = x
    a.cutoff_entry.set("250000")
    a.filter_data()
    want = to.filter_data(x.astype(np.float32).astype(np.float64), a.samplerate, 250000.0, 8)
    assert np.abs(a.filtered_data.cpu().numpy() - want).max() <= FILTER_TOL_PA


@pytest.mark.parametrize("normalize", [0, 1])
def test_update_psd_default_length(tmp_path, normalize):
    """2.1 M samples: nperseg = 2^20, three segments (plot-trace.py:437)."""
    paths = write_series(tmp_path)
    a = ctapp.App(None, paths[0])
    a.start_entry.set("0")
    a.end_entry.set("10")                          # beyond the series: clamped to total_samples
    a.cutoff_entry.set("100000")
    a.normalize.set(normalize)
    a.update_psd()
    data, fs = to.load_mapped_data(paths[0], 0.0, 10.0)
    assert len(data) == 2_100_000
    filt = chain_input(a, to.filter_data(data, fs, 100000.0, 8))
    f, P, rms, cur = to.update_psd(filt, fs, None, bool(normalize), 100000.0)
    assert np.allclose(a.f, f, rtol=1e-14) and psd_close(a.Pxx, P)
    assert np.allclose(a.rms, rms, rtol=1e-6) and np.isclose(a.current, cur, rtol=1e-6)
    assert a.psd_limits[:2] == (1, 200000.0)
    lo = np.log10(P[1:np.searchsorted(f, 100000.0)])
    assert a.psd_limits[2] == 10 ** np.floor(lo.min()) and a.psd_limits[3] == 10 ** np.ceil(lo.max())
    N = len(f[f < 10000])
    want_norm = P[1:N] if normalize else P[1:N] * 100000.0 / cur ** 2
    assert np.allclose(a.fnorm, f[1:N]) and np.allclose(a.Pxx_norm, want_norm, rtol=1e-4)
    a.export_psd(str(tmp_path / "psd.csv"))
    back = np.loadtxt(tmp_path / "psd.csv", delimiter=",")
    assert back.shape == (len(f), 3) and np.array_equal(back[:, 0], a.f) and np.array_equal(back[:, 2], a.rms)


def test_update_psd_with_a_length_entry_and_no_filter(tmp_path):
    """psd_length_entry in seconds -> 2**ceil(log2(.)) (a float, plot-trace.py:433); no cutoff: bandwidth 1 MHz."""
    paths = write_series(tmp_path)
    a = ctapp.App(None, paths[0])
    a.end_entry.set("0.3")
    a.cutoff_entry.set("")
    a.psd_length_entry.set("0.01")                # 41 667 samples -> 65 536
    a.normalize.set(1)
    a.update_psd()
    data, fs = to.load_mapped_data(paths[0], 0.0, 0.3)
    f, P, rms, cur = to.update_psd(data.astype(np.float32).astype(np.float64), fs, 0.01, True, None)
    assert len(a.f) == 32769 and np.allclose(a.f, f, rtol=1e-14) and psd_close(a.Pxx, P)
    assert np.allclose(a.rms, rms, rtol=1e-6) and a.psd_limits[1] == 2e6


def test_update_psd_on_a_window_shorter_than_the_segment(tmp_path):
    """length > len(data) -> nperseg = len(data): one segment of arbitrary length (plot-trace.py:434-435)."""
    paths = write_series(tmp_path)
    a = ctapp.App(None, paths[0])
    a.end_entry.set("0.05")                        # 208 333 samples
    a.cutoff_entry.set("100000")
    a.psd_length_entry.set("0.1")
    a.update_psd()
    data, fs = to.load_mapped_data(paths[0], 0.0, 0.05)
    filt = chain_input(a, to.filter_data(data, fs, 100000.0, 8))
    f, P, rms, cur = to.update_psd(filt, fs, 0.1, False, 100000.0)
    assert len(a.f) == len(data) // 2 + 1 and np.allclose(a.f, f) and psd_close(a.Pxx, P, single=True)
    assert np.allclose(a.rms, rms, rtol=1e-5)


def test_scale_raw_data_accepts_numpy_and_reports_a_rate_mismatch(tmp_path):
    paths = write_series(tmp_path, n_each=20_000)
    a = ctapp.App(None, paths[0])
    codes = np.fromfile(paths[0], dtype=np.uint16)
    want = to.scale_raw_data(codes, synth.CHIMERA_SETTINGS).astype(np.float32)
    got = a.scale_raw_data(codes, a.settings[0]).cpu().numpy()
    assert np.all(np.abs(got - want) <= np.spacing(np.abs(want))) and np.mean(got == want) > 0.9999
    assert a.scale_raw_data(codes[:0], a.settings[0]).numel() == 0
    other = dict(a.settings[0])
    other["ADCSAMPLERATE"] = 1.0e6
    a.scale_raw_data(codes[:10], other)
    assert a.wildcard.get() == "One of your files does not match the global sampling rate!"


def test_legacy_psd_app(tmp_path):
    """legacy/minimal_psd.py: (>i2, >i2) records, savegain, filter, 2^18-point Welch, normalisation by maxf / 2."""
    rng = np.random.default_rng(2)
    n = 700_000
    rec = np.zeros(n, dtype=np.dtype([("current", ">i2"), ("voltage", ">i2")]))
    rec["current"] = np.clip(np.rint(8000 + 300 * rng.standard_normal(n)), -32768, 32767)
    rec["voltage"] = 3
    p = tmp_path / "legacy.dat"
    rec.tofile(p)
    a = ctapp.LegacyPsdApp(None, str(p))
    a.samplerate_entry.set("500000")
    a.savegain_entry.set("0.25")
    a.start_entry.set("0.1")
    a.end_entry.set("5")                           # clamped to the file
    a.cutoff_entry.set("50000")
    a.order_entry.set("8")
    a.normalize.set(1)
    a.update_psd()
    data = to.load_legacy_i2(str(p), 50_000, n, 0.25)
    assert np.array_equal(a.data.cpu().numpy(), data.astype(np.float32))
    filt = chain_input(a, to.filter_data(data, 500000.0, 50000.0, 8))
    L = 2 ** 18
    f, P = to.welch_psd(filt, 500000.0, L)
    cur = np.average(filt[:len(filt) // L * L])
    assert np.allclose(a.f, f, rtol=1e-14) and np.allclose(a.rms, to.integrate_noise(f, P), rtol=1e-6)
    assert psd_close(a.Pxx, P / cur ** 2 * 50000.0)
    a.cutoff_entry.set("")
    a.update_trace()
    assert a.plot_data is a.data
    a.end_entry.set("")
    with pytest.raises(UnboundLocalError):         # legacy/minimal_psd.py:188 with an empty end entry
        a.load_mapped_data()


@pytest.mark.parametrize("poles", [2, 4, 8])
def test_bessel_step_app(poles):
    """legacy/bessel-filter.py:100-131 with the tool's own entries (kHz): 5 fs / fc samples of a unit step."""
    a = ctapp.BesselStepApp()
    a.fc_entry.set("100")
    a.fs_entry.set("4166.666")
    a.poles.set(str(poles))
    a.update_filter()
    n = int(5 * 4166666.0 / 100000.0)
    step = np.zeros(n)
    step[n // 2:] = 1
    assert np.array_equal(a.perfect_data.cpu().numpy(), step.astype(np.float32))
    want = to.filter_data_edge(step, 4166666.0, 100000.0, poles)
    got = a.filtered_data.cpu().numpy()
    assert got.shape == want.shape and np.abs(got - want).max() <= 2e-6


def test_spectrum_sample_class(tmp_path, golden_dir):
    """noise-fit.py:84-100 on the committed fixture (outputs of the reference's own class) and its curve fit."""
    z = np.load(os.path.join(golden_dir, "spectrum_fixture.npz"))
    raw = z["raw"]
    rec = np.zeros(len(raw), dtype=np.dtype([("curr_pA", ">f8"), ("volt_mV", ">f8")]))
    rec["curr_pA"] = raw
    p = tmp_path / "B0001.bin"
    rec.tofile(p)
    psdlength = 2 ** np.ceil(np.log2(float(z["psdlength"])))
    s = ctapp.SpectrumSample(str(p), float(z["fs"]), psdlength, float(z["cutoff"]))
    f, P, cur = to.spectrum_sample(raw, float(z["fs"]), psdlength, float(z["cutoff"]))
    assert np.allclose(s.f, f) and np.allclose(s.Pxx, P, rtol=2e-5) and np.isclose(s.current, cur, rtol=1e-6)
    assert np.allclose(s.Pxx, z["Pxx"], rtol=2e-5) and np.array_equal(s.f, z["f"])
    assert (s.thermal, s.pink, s.brown) == (1.0e-3, 1, 1.0e-3)
    s.fit_spectrum()
    assert np.all(np.isfinite([s.thermal, s.pink, s.brown])) and len(s.p0) == 3


def test_print_trace(tmp_path):
    rng = np.random.default_rng(4)
    n = 50_000
    rec = np.zeros(n, dtype=np.dtype([("curr_pA", ">f8"), ("volt_mV", ">f8")]))
    rec["curr_pA"] = 5000 + 100 * rng.standard_normal(n)
    p = tmp_path / "trace.bin"
    rec.tofile(p)
    out = ctapp.print_trace(str(p), 0.002, 0.005, 4166666)
    assert out == str(tmp_path / "trace") + "_0.002_0.005_current.csv"
    want = to.load_bin(str(p), 0.002, 0.005, 4166666)
    got = np.loadtxt(out)
    assert got.shape == want.shape and np.array_equal(got, want.astype(np.float32).astype(np.float64))
