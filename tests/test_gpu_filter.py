"""GPU parity: fused dequantise + median-pad + Bessel filtfilt kernel vs the CPU oracle
(the reference's own call sequence) and vs fixtures produced by the reference itself.

Tolerances (float32 state with median subtraction, Wn >= 0.04):
  |gpu - ref| <= 0.05 pA  and  <= 1e-5 of the baseline level.
The reference's float64 direct form has its own rounding-noise floor of ~3e-4 pA."""
import os

import numpy as np
import pytest
import torch
from scipy.signal import bessel, lfilter, lfilter_zi

from cusumtools_b200 import filters, synth
from cusumtools_b200.design import bessel_lowpass
from oracle import trace_oracle as to

pytestmark = pytest.mark.gpu
ABS_TOL = 0.05
S = synth.CHIMERA_SETTINGS


def gpu_filter(codes, cutoff, order, **kw):
    raw = torch.from_numpy(codes).cuda()
    y = filters.dequant_filtfilt(raw, S, cutoff, order, **kw)
    torch.cuda.synchronize()
    return y.cpu().numpy()


def check(got, want):
    err = np.abs(got.astype(np.float64) - want)
    assert err.max() <= ABS_TOL, err.max()
    assert err.max() <= 1e-5 * np.abs(want).max(), err.max()


def test_c1_slice_matches_oracle():
    codes, _ = synth.c1_trace(n=400000, n_events=95, seed=0)
    want = to.filter_data(to.scale_raw_data(codes, S), synth.FS, 1e5, 8)
    check(gpu_filter(codes, 1e5, 8), want)


@pytest.mark.parametrize("key", ["filt_100000_8", "filt_250000_8", "filt_900000_8", "filt_100000_4", "filt_200000_5"])
def test_reference_fixture(golden_dir, key):
    z = np.load(os.path.join(golden_dir, "filter_fixture.npz"))
    _, cutoff, order = key.split("_")
    check(gpu_filter(z["codes"], float(cutoff), int(order)), z[key])


def test_reference_fixture_even_length(golden_dir):
    z = np.load(os.path.join(golden_dir, "filter_fixture.npz"))
    check(gpu_filter(z["codes"][:-1].copy(), 1e5, 8), z["filt_even_100000_8"])


@pytest.mark.parametrize("n", [1, 2, 17, 511, 512, 513, 1000, 4095, 4096, 4097, 9000])
def test_ragged_lengths(n):
    codes, _ = synth.c1_trace(n=max(n, 3000), n_events=1, seed=n)
    codes = codes[:n].copy()
    want = to.filter_data(to.scale_raw_data(codes, S), synth.FS, 2.5e5, 8)
    check(gpu_filter(codes, 2.5e5, 8), want)


def test_empty_trace_is_rejected():
    raw = torch.zeros(0, dtype=torch.uint16, device="cuda")
    with pytest.raises(ValueError):
        filters.dequant_filtfilt(raw, S, 1e5, 8)


def test_unaligned_views():
    codes, _ = synth.c1_trace(n=70001, n_events=15, seed=3)
    raw = torch.from_numpy(codes).cuda()
    view = raw[1:]                       # 2-byte offset: vector loads are not possible
    out = torch.empty(70001, dtype=torch.float32, device="cuda")[1:]
    y = filters.dequant_filtfilt(view, S, 1e5, 8, out=out)
    torch.cuda.synchronize()
    want = to.filter_data(to.scale_raw_data(codes[1:], S), synth.FS, 1e5, 8)
    check(y.cpu().numpy(), want)


def test_halo_sharding_is_invisible():
    """Sub-segments are filtered independently with an H-sample warm-up: the result must
    not depend on the sub-segment length (512 vs 4096 vs 8192)."""
    codes, _ = synth.c1_trace(n=300000, n_events=70, seed=9)
    a = gpu_filter(codes, 1e5, 8, subsegment=512)
    b = gpu_filter(codes, 1e5, 8, subsegment=4096)
    c = gpu_filter(codes, 1e5, 8, subsegment=8192)
    assert np.abs(a - b).max() < 0.02 and np.abs(b - c).max() < 0.02


def test_linearity_and_dc_gain():
    # constant input -> constant output (DC gain exactly 1, steady-state boundaries)
    codes = np.full(50000, 40000, np.uint16) & np.uint16(filters.chimera_bitmask(S))
    y = gpu_filter(codes, 1e5, 8)
    assert np.ptp(y) == 0 and np.isclose(y[0], to.scale_raw_data(codes[:1], S)[0], rtol=1e-7)


def test_forward_only_matches_lfilter():
    codes, _ = synth.c1_trace(n=100000, n_events=20, seed=4)
    x = to.scale_raw_data(codes, S)
    b, a = bessel(8, 2 * 1e5 / synth.FS, "low")
    med = np.median(x)
    want, _ = lfilter(b, a, x, zi=lfilter_zi(b, a) * med)
    got = gpu_filter(codes, 1e5, 8, forward_only=True, padding=0)
    check(got, want)


def test_float_input_path():
    codes, _ = synth.c1_trace(n=120000, n_events=25, seed=6)
    x = to.scale_raw_data(codes, S)
    med = float(np.median(x))
    want = to.filter_data(x, synth.FS, 1e5, 8)
    xt = torch.from_numpy(x.astype(np.float32)).cuda()
    got = filters.bessel_filtfilt(xt, synth.FS, 1e5, 8, pad_value=med).cpu().numpy()
    check(got, want)
    # plain zero-state lfilter
    b, a = bessel(8, 2 * 1e5 / synth.FS, "low")
    want2 = lfilter(b, a, x - med)
    got2 = filters.bessel_lfilter(torch.from_numpy((x - med).astype(np.float32)).cuda(), synth.FS, 1e5, 8).cpu().numpy()
    assert np.abs(got2 - want2).max() < ABS_TOL


def test_code_median_is_exact():
    rng = np.random.default_rng(11)
    for n in (1, 2, 1001, 65536, 5_000_001):
        codes = (rng.normal(40000, 60, n).astype(np.int64) & 0xFFFC).astype(np.uint16)
        c1, c2 = filters.code_median(torch.from_numpy(codes).cuda(), 0xFFFC)
        srt = np.sort(codes)
        assert (c1, c2) == (int(srt[(n - 1) // 2]), int(srt[n // 2]))
    # bimodal: the window search must recover when the sample estimate is off
    codes = np.concatenate((np.full(500000, 1000), np.full(500001, 60000))).astype(np.uint16)
    assert filters.code_median(torch.from_numpy(codes).cuda(), 0xFFFF) == (60000, 60000)


def test_large_trace_properties():
    """2^27 samples generated on the device: time-reversal symmetry of the zero-phase
    filter (filtfilt(reverse(x)) == reverse(filtfilt(x))) and agreement with the oracle on
    an interior window."""
    n = 1 << 27
    raw = synth.device_trace(n, "cuda", seed=99)
    c1, c2 = filters.code_median(raw, filters.chimera_bitmask(S))
    y = filters.dequant_filtfilt(raw, S, 1e5, 8, median_codes=(c1, c2))
    yr = filters.dequant_filtfilt(torch.flip(raw.view(torch.int16), [0]).view(torch.uint16).contiguous(), S, 1e5, 8, median_codes=(c1, c2))
    assert torch.max(torch.abs(torch.flip(yr, [0]) - y)).item() < 0.05
    lo, hi = 50_000_000, 50_300_000
    win = raw[lo - 5000:hi + 5000].cpu().numpy()
    x = to.scale_raw_data(win, S)
    want = to.filter_data(x, synth.FS, 1e5, 8)[5000:-5000]
    got = y[lo:hi].cpu().numpy()
    assert np.abs(got - want).max() < ABS_TOL


def test_lane_sequential_and_warp_scan_paths_agree():
    """The default path (two lane-sequential passes) and the single-kernel warp-scan path
    compute the same cascade with different association: they agree to float32 noise."""
    codes, _ = synth.c1_trace(n=700000, n_events=170, seed=21)
    a = gpu_filter(codes, 1e5, 8)
    b = gpu_filter(codes, 1e5, 8, subsegment=4096)
    assert np.abs(a - b).max() < 0.02
