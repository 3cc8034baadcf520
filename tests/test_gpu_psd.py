"""GPU parity: hand-written FFT Welch vs scipy.signal.welch as the reference calls it.
Tolerance: rel 5e-6 per bin for a float32 FFT, stated against the
spectrum's own dynamic range: |dP| <= 5e-6 P + 1e-9 max(P) (bins more than ~9 decades below
the peak are below the float32 transform's noise floor); cumulative rms rel <= 1e-6."""
import os

import numpy as np
import pytest
import torch

from cusumtools_b200 import psd, synth
from oracle import trace_oracle as to

pytestmark = pytest.mark.gpu


def check(P, want):
    tol = 5e-6 * want + 1e-9 * want.max()
    bad = np.abs(P - want) > tol
    assert not bad.any(), (np.nonzero(bad)[0][:10], (np.abs(P - want) / tol).max())


@pytest.mark.parametrize("L", [256, 512, 1024, 4096, 1 << 14, 1 << 15, 1 << 16, 1 << 17, 1 << 18, 1 << 19])
def test_white_noise_all_sizes(L):
    rng = np.random.default_rng(L)
    x = (5000 + 150 * rng.standard_normal(5 * L + 123)).astype(np.float32)
    f, P = psd.welch(torch.from_numpy(x).cuda(), synth.FS, L)
    fw, Pw = to.welch_psd(x.astype(np.float64), synth.FS, L)
    assert np.allclose(f, fw, rtol=1e-14)
    check(P, Pw)
    assert np.allclose(psd.integrate_noise(f, P), to.integrate_noise(fw, Pw), rtol=1e-6)


@pytest.mark.parametrize("L", [1 << 21, 1 << 22, 1 << 23])
def test_long_segments(L):
    """Both kernel families at their largest sizes (register-resident FFT up to 2^22, shared-memory
    Stockham beyond), 4 segments, with |x| as noise-fit.py:92 passes it."""
    rng = np.random.default_rng(L + 1)
    x = (40 * rng.standard_normal(5 * L // 2 + 17)).astype(np.float32)
    f, P = psd.welch(torch.from_numpy(x).cuda(), 5.0e5, L, use_abs=True)
    fw, Pw = to.welch_psd(np.abs(x.astype(np.float64)), 5.0e5, L)
    check(P, Pw)


@pytest.mark.parametrize("n", [17, 255, 1000, 4097, 65536 + 1, 300_001, 1_000_000, 999_983, (1 << 20) - 1, 1_500_000])
def test_single_segment_of_arbitrary_length(n):
    """nperseg == len(x), not a power of two: what plot-trace.py:433-437 passes for a window shorter than 2^20
    samples (or than the requested PSD length).  Chirp-z over the float32 FFT: |dP| <= 5e-5 P + 1e-8 max(P)
    (three transforms instead of one; 999 983 is prime)."""
    rng = np.random.default_rng(n)
    t = np.arange(n)
    x = (5000 + 24 * rng.standard_normal(n) + 40 * np.sin(2 * np.pi * 0.01 * t)).astype(np.float32)
    f, P = psd.welch(torch.from_numpy(x).cuda(), synth.FS, n)
    fw, Pw = to.welch_psd(x.astype(np.float64), synth.FS, n)
    assert f.shape == fw.shape and np.allclose(f, fw, rtol=1e-14)
    tol = 5e-5 * Pw + 1e-8 * Pw.max()
    bad = np.abs(P - Pw) > tol
    assert not bad.any(), (np.nonzero(bad)[0][:10], (np.abs(P - Pw) / tol).max())
    assert np.allclose(psd.integrate_noise(f, P), to.integrate_noise(fw, Pw), rtol=1e-5)


def test_update_psd_on_a_short_window():
    """App.update_psd on 0.1 s of data (416 666 samples < 2^20): length = len(data), plot-trace.py:437."""
    rng = np.random.default_rng(8)
    x = (5000 + 24 * rng.standard_normal(416_666)).astype(np.float32)
    f, P, rms, cur = psd.update_psd(torch.from_numpy(x).cuda(), synth.FS)
    fw, Pw = to.welch_psd(x.astype(np.float64), synth.FS, len(x))
    assert len(f) == len(x) // 2 + 1 and np.allclose(f, fw)
    assert np.allclose(rms, to.integrate_noise(fw, Pw), rtol=1e-5) and np.isclose(cur, x.astype(np.float64).mean(), rtol=1e-9)


def test_reference_fixture(golden_dir):
    z = np.load(os.path.join(golden_dir, "psd_fixture.npz"))
    x = z["x"].astype(np.float32)
    f, P = psd.welch(torch.from_numpy(x).cuda(), float(z["fs"]), int(z["nperseg"]))
    fw, Pw = to.welch_psd(x.astype(np.float64), float(z["fs"]), int(z["nperseg"]))
    check(P, Pw)
    # against the fixture itself (float64 input there): dominated by the float32 cast of x
    assert np.allclose(P[:200], z["Pxx"][:200], rtol=1e-4)
    assert np.allclose(psd.integrate_noise(f, P), z["rms"], rtol=1e-5)


def test_sinusoid_and_dc():
    L = 1 << 14
    n = 8 * L
    t = np.arange(n)
    x = (1000 + 50 * np.sin(2 * np.pi * 37.25 * t / L) + 3 * np.cos(2 * np.pi * 0.5 * t / L)).astype(np.float32)
    f, P = psd.welch(torch.from_numpy(x).cuda(), 1.0e6, L)
    fw, Pw = to.welch_psd(x.astype(np.float64), 1.0e6, L)
    check(P, Pw)


def test_full_size_segments():
    """The reference's default 2^20-point segments (config C4 shape, 7 segments)."""
    L = 1 << 20
    rng = np.random.default_rng(4)
    x = (5000 + 24 * rng.standard_normal(4 * L)).astype(np.float32)
    f, P = psd.welch(torch.from_numpy(x).cuda(), synth.FS, L)
    fw, Pw = to.welch_psd(x.astype(np.float64), synth.FS, L)
    check(P, Pw)
    rms = psd.integrate_noise(f, P)
    assert abs(rms[-1] - x.astype(np.float64).std()) / rms[-1] < 2e-3     # Parseval


def test_update_psd_and_spectrum_sample(golden_dir):
    z = np.load(os.path.join(golden_dir, "spectrum_fixture.npz"))
    raw = z["raw"].astype(np.float32)
    f, P, cur = psd.spectrum_sample(torch.from_numpy(raw).cuda(), float(z["fs"]), 2 ** np.ceil(np.log2(8192)), float(z["cutoff"]))
    assert np.allclose(f, z["f"]) and np.isclose(cur, z["current"], rtol=1e-6)
    assert np.allclose(P, z["Pxx"], rtol=2e-4)
    rng = np.random.default_rng(2)
    y = (5000 + 24 * rng.standard_normal(3 * (1 << 16) + 99)).astype(np.float32)
    got = psd.update_psd(torch.from_numpy(y).cuda(), synth.FS, psd_length_s=0.012, normalize=True, cutoff=1e5)
    want = to.update_psd(y.astype(np.float64), synth.FS, 0.012, True, 1e5)
    assert np.allclose(got[1], want[1], rtol=1e-4, atol=1e-9 * want[1].max()) and np.allclose(got[2], want[2], rtol=1e-5)


def test_unsupported_lengths_fail_loudly():
    x = torch.zeros(5000, dtype=torch.float32, device="cuda")
    with pytest.raises(NotImplementedError):
        psd.welch(x, 1e6, 3000)      # neither a power of two nor the whole input (the reference never asks for it)
    with pytest.raises(NotImplementedError):
        psd.welch(torch.zeros((1 << 21) + 5, dtype=torch.float32, device="cuda"), 1e6, (1 << 21) + 5)
    f, P = None, None
    with pytest.raises(ValueError):
        psd.welch(x, 1e6, 8192)      # shorter than one segment
