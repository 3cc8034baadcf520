"""Stage 4 of the hot path: Welch PSD and the noise integrals built on it.

Mirrors `scipy.signal.welch(x, fs, nperseg=L)` exactly as the reference calls it
(plot-trace.py:432-448, noise-fit.py:92-99, legacy/minimal_psd.py:250-255) plus
`App.integrate_noise` (plot-trace.py:309-311).  The periodogram sums come from the
hand-written FFT kernels (ct_welch_f32); the O(L) scaling / cumulative sum runs on the host
in float64 like the reference's."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .filters import _require_cuda, _stream_ptr

L2_BUDGET_BYTES = 1 << 30      # four-step intermediate of one batch of segments (bigger batches measured faster up to
                               # ~1 GB: more segments per CTA - C4 16.9 ms at 512 MB, 15.9 ms at 1 GB; L2 residency of the
                               # intermediate does not matter; two half-batches on two streams: 16.2 ms, rejected)


def psd_length(n: int, samplerate: float, psd_length_s: float | None = None) -> int:
    """plot-trace.py:432-437: segment length from the GUI entry (seconds) or min(2^20, n)."""
    if psd_length_s is not None:
        length = 2 ** np.ceil(np.log2(float(psd_length_s) * samplerate))
        if length > n:
            length = n
    else:
        length = np.minimum(2 ** 20, n)
    return int(length)


def welch_sums(x: torch.Tensor, nperseg: int, *, use_abs: bool = False, shift: float | None = None):
    """Raw periodogram sums on the device: (acc float64[L/2+1] CUDA tensor, nseg).  A
    multi-GPU caller reduces `acc` and `nseg` over ranks before scaling."""
    _require_cuda(x, "x", torch.float32)
    L = int(nperseg)
    if L != nperseg or L & (L - 1) or L < 256:
        raise NotImplementedError(
            f"nperseg={nperseg}: the GPU Welch path needs a power-of-two segment >= 256 (the reference's default "
            "2**20 and its 2**ceil(log2(.)) lengths are) or nperseg == len(x) (one segment of arbitrary length)")
    lib = _lib.lib()
    batch = max(1, min(1024, L2_BUDGET_BYTES // (L // 2 * 8)))
    wsb = int(lib.ct_welch_workspace_bytes(L, batch))
    ws = torch.empty(wsb, dtype=torch.uint8, device=x.device)
    acc = torch.empty(L // 2 + 1, dtype=torch.float64, device=x.device)
    if shift is None:
        shift = float(x[: min(x.numel(), 1 << 16)].abs().mean().item() if use_abs else x[: min(x.numel(), 1 << 16)].mean().item()) if x.numel() else 0.0
    nseg = C.c_int64(0)
    with torch.cuda.device(x.device):   # the library launches on the current device
        rc = lib.ct_welch_f32(x.data_ptr(), x.numel(), L, float(shift), int(bool(use_abs)), batch, ws.data_ptr(), wsb,
                              acc.data_ptr(), C.byref(nseg), _stream_ptr(x))
    _lib.check(rc, "ct_welch_f32")
    return acc, int(nseg.value)


def scale_sums(acc: np.ndarray, nseg: int, samplerate: float, nperseg: int):
    """Density scaling of scipy's welch: / (fs * sum(w^2)), mean over segments, one-sided
    doubling of bins 1..L/2-1.  sum(w^2) = 3L/8 exactly for the periodic Hann window."""
    L = int(nperseg)
    if nseg <= 0:
        raise ValueError("trace shorter than one segment")
    P = np.asarray(acc, dtype=np.float64) / (float(nseg) * float(samplerate) * (3.0 * L / 8.0))
    P[1:-1] *= 2.0
    f = np.arange(L // 2 + 1, dtype=np.float64) * (float(samplerate) / L)
    return f, P


def welch_single_sums(x: torch.Tensor, *, use_abs: bool = False) -> torch.Tensor:
    """|rfft(hann(n) (x - mean x))|^2 of ONE segment of arbitrary length n = len(x) (float64 CUDA tensor
    [n//2+1]): the case nperseg == len(x) of plot-trace.py:433-437 when the window is not a power of two.
    Bluestein chirp-z over the power-of-two FFT kernels (ct_welch_single_f32)."""
    _require_cuda(x, "x", torch.float32)
    n = int(x.numel())
    lib = _lib.lib()
    wsb = int(lib.ct_welch_single_workspace_bytes(n))
    if wsb < 0:
        raise NotImplementedError(f"a single Welch segment of {n} samples: the chirp-z path covers 2 <= n <= 2**21")
    ws = torch.empty(wsb, dtype=torch.uint8, device=x.device)
    acc = torch.empty(n // 2 + 1, dtype=torch.float64, device=x.device)
    xm = x.abs() if use_abs else x
    mean = float(xm.sum(dtype=torch.float64).item()) / n         # scipy's detrend='constant' of the segment
    with torch.cuda.device(x.device):
        rc = lib.ct_welch_single_f32(x.data_ptr(), n, mean, int(bool(use_abs)), ws.data_ptr(), wsb, acc.data_ptr(), _stream_ptr(x))
    _lib.check(rc, "ct_welch_single_f32")
    return acc


def welch(x: torch.Tensor, samplerate: float, nperseg, *, use_abs: bool = False):
    """`f, Pxx = welch(x, fs, nperseg=L)` with x a float32 CUDA tensor; returns numpy
    float64 arrays of length L//2+1 like scipy.  `nperseg` may be a float (the reference
    passes 2**np.ceil(...), plot-trace.py:433, noise-fit.py:138).  L is a power of two, or the
    length of `x` itself (plot-trace.py:435,437: one segment of arbitrary length)."""
    L = int(nperseg)
    if (L & (L - 1) or L < 256) and L == x.numel() and L >= 16:
        acc = welch_single_sums(x, use_abs=use_abs).cpu().numpy()
        P = acc / (float(samplerate) * (3.0 * L / 8.0))          # sum(w^2) = 3L/8 for the periodic Hann, one segment
        if L % 2 == 0:
            P[1:-1] *= 2.0
        else:
            P[1:] *= 2.0
        return np.arange(L // 2 + 1, dtype=np.float64) * (float(samplerate) / L), P
    acc, nseg = welch_sums(x, L, use_abs=use_abs)
    return scale_sums(acc.cpu().numpy(), nseg, samplerate, L)


def integrate_noise(f: np.ndarray, Pxx: np.ndarray) -> np.ndarray:
    """plot-trace.py:309-311."""
    df = f[1] - f[0]
    return np.sqrt(np.cumsum(Pxx * df))


def update_psd(filtered: torch.Tensor, samplerate: float, *, psd_length_s: float | None = None,
               normalize: bool = False, cutoff: float | None = None):
    """`App.update_psd` without the plotting (plot-trace.py:418-451): returns
    (f, Pxx, rms, current).  `cutoff` None means the trace was not filtered (bandwidth 1 MHz)."""
    _require_cuda(filtered, "filtered", torch.float32)
    n = filtered.numel()
    bandwidth = 1.0e6 if cutoff is None else float(cutoff)
    length = psd_length(n, samplerate, psd_length_s)
    end_index = int(np.floor(n / length) * length)
    current = float(filtered[:end_index].to(torch.float64).mean().item())
    f, Pxx = welch(filtered, samplerate, length)
    rms = integrate_noise(f, Pxx)
    if normalize:
        Pxx = Pxx / current ** 2 * bandwidth
        rms = rms / np.absolute(current)
    return f, Pxx, rms, current


def spectrum_sample(raw: torch.Tensor, samplerate: float, psdlength, cutoff: float):
    """`SpectrumSample.__init__` (noise-fit.py:92-99): welch(|raw|), crop to f <= cutoff,
    scale by cutoff / I^2 and by f.  Returns (f, Pxx, current)."""
    _require_cuda(raw, "raw", torch.float32)
    f, Pxx = welch(raw, samplerate, psdlength, use_abs=True)
    inds = f <= cutoff
    length = int(np.sum(inds)) - 1
    f = f[1:length]
    Pxx = Pxx[1:length].copy()
    current = abs(float(raw.to(torch.float64).mean().item()))
    Pxx *= cutoff / current ** 2
    Pxx *= f
    return f, Pxx, current


# ---------------------------------------------------------------- exports
def export_psd(path: str, f: np.ndarray, Pxx: np.ndarray, rms: np.ndarray) -> None:
    """`App.export_psd` (plot-trace.py:204-207): rows `f,Pxx,rms`, numpy's default '%.18e'."""
    np.savetxt(path, np.c_[f, Pxx, rms], delimiter=",")


def export_psd_tsv(path: str, f: np.ndarray, Pxx: np.ndarray, current: float, bandwidth: float) -> None:
    """The 4-column tab-separated `.psd` file `legacy/psdfit.py:27` reads
    (`names=['f','S','integral','norm']`, no header): frequency, PSD, cumulative integral
    of the PSD (the square of `integrate_noise`), and the PSD normalised as
    plot-trace.py:445-447 does (`S * bandwidth / I^2`)."""
    df = f[1] - f[0]
    integral = np.cumsum(Pxx * df)
    norm = Pxx / float(current) ** 2 * float(bandwidth)
    np.savetxt(path, np.c_[f, Pxx, integral, norm], delimiter="\t")


# ------------------------------------------------------------------ multi-GPU
def segment_share(n: int, nperseg: int, world: int, rank: int) -> tuple[int, int, int, int]:
    """Contiguous block of Welch segments owned by `rank` and the samples it must hold:
    (first_segment, last_segment_exclusive, first_sample, last_sample_exclusive).  Segment s
    covers samples [s*hop, s*hop + L), hop = L/2, s < (n - L/2) // hop (scipy's count)."""
    L = int(nperseg)
    hop = L // 2
    nseg = max(0, (int(n) - (L - hop)) // hop)
    per = -(-nseg // int(world))
    s0, s1 = min(nseg, rank * per), min(nseg, (rank + 1) * per)
    return s0, s1, s0 * hop, (s1 - 1) * hop + L if s1 > s0 else s0 * hop


def reduce_sums(acc: torch.Tensor, nseg: int, group=None):
    """Sum the per-rank periodogram sums and segment counts over `group`: the one collective of the PSD stage,
    L/2+1 float64 values with the integer count riding as one more (exact below 2^53); returns (acc numpy, nseg)."""
    if group is not None:
        import torch.distributed as dist
        packed = torch.cat((acc, torch.tensor([float(int(nseg))], dtype=torch.float64, device=acc.device)))
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
        acc, nseg = packed[:-1], int(round(float(packed[-1].item())))
    return acc.cpu().numpy(), int(nseg)


def welch_sharded(x_local: torch.Tensor, samplerate: float, nperseg, *, use_abs: bool = False, shift: float | None = None,
                  group=None):
    """`welch` over a trace whose segments are split over the ranks of `group`: `x_local`
    holds exactly the samples `segment_share` assigns to this rank.  Every rank must pass
    the same `shift` (any constant near the global mean; None = 0 for sharded calls would
    lose float32 headroom, so the caller passes the pad value / global baseline)."""
    L = int(nperseg)
    if x_local.numel() >= L:
        acc, nseg = welch_sums(x_local, L, use_abs=use_abs, shift=shift)
    else:
        acc, nseg = torch.zeros(L // 2 + 1, dtype=torch.float64, device=x_local.device), 0
    a, n = reduce_sums(acc, nseg, group)
    return scale_sums(a, n, samplerate, L)
