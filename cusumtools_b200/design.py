"""Host-side low-pass Bessel design for the fused dequantise + filtfilt kernel.

Replaces the `bessel(order, Wn, 'low')` call at reference `plot-trace.py:314-317`
(and `legacy/minimal_psd.py:199-203`, `legacy/bessel-filter.py:124-128`).  The
reference obtains direct-form `b, a`; a float32 kernel cannot run that form (the
rounded 9-tap polynomial is unstable), so this module
derives the same filter as poles -> second-order all-pole sections, all in
float64 on the host.  It is a few microseconds of scalar work, not hot path.

Mathematics (phase-normalised Bessel, the scipy default the reference relies on):
  * analog prototype poles = reciprocals of the zeros of the ordinary Bessel
    polynomial y_N, scaled by a_last**(-1/N) with a_last = (2N)!/(N! 2^N);
  * low-pass frequency scaling by the pre-warped `4 tan(pi Wn / 2)`;
  * bilinear transform at fs = 2:  p_d = (4 + p) / (4 - p);
  * every zero of the digital filter sits at z = -1, i.e. the numerator is
    g * (1 + z^-1)^N exactly.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

MAX_SECTIONS = 5  # orders 1..10


def _reverse_bessel_coeffs(n: int) -> list[int]:
    """Exact integer coefficients of the reverse Bessel polynomial theta_n(s),
    highest power first: [1, 1], [1, 3, 3], [1, 6, 15, 15], ..."""
    out = []
    for k in range(n + 1):
        num = 1
        for j in range(n - k + 1, 2 * n - k + 1):
            num *= j
        out.append(num // (2 ** (n - k) * math.factorial(k)))
    return out[::-1]


def bessel_analog_poles(order: int) -> np.ndarray:
    """Poles of the phase-normalised analog Bessel prototype (complex128).

    Roots of theta_n are found with numpy's companion-matrix solver and polished
    with Newton steps on the exact integer polynomial, then conjugate pairs are
    symmetrised."""
    if order < 1 or order > 2 * MAX_SECTIONS:
        raise ValueError(f"Bessel order must be in 1..{2 * MAX_SECTIONS}, got {order}")
    c = _reverse_bessel_coeffs(order)
    cf = np.array(c, dtype=np.float64)
    dcf = np.polyder(cf)
    r = np.roots(cf).astype(np.complex128)
    for _ in range(8):
        r = r - np.polyval(cf, r) / np.polyval(dcf, r)
    # order by imaginary part so conjugates pair up (a real root lands in the middle)
    r = r[np.argsort(r.imag)]
    r = 0.5 * (r + np.conj(r[::-1]))
    a_last = c[-1]
    return r * 10 ** (-math.log10(a_last) / order)


@dataclass(frozen=True)
class BesselDesign:
    """A digital low-pass Bessel filter as the kernels consume it.

    sections[i] = (a1, a2): all-pole denominator 1 + a1 z^-1 + a2 z^-2 (a2 == 0
    and `first_order[i]` True for the real pole of an odd order).  The numerator
    of section i is (1 + z^-1)^2 (or (1 + z^-1) for a first-order section) and
    `gain` is the overall scale that makes the DC gain exactly 1."""
    order: int
    wn: float
    poles: np.ndarray            # digital poles, complex128, all N of them
    sections: np.ndarray         # float64 [nsec, 2]
    first_order: np.ndarray      # bool [nsec]
    gain: float                  # product over sections of (1+a1+a2)/4 (or /2)
    section_gain: np.ndarray     # per-section DC normaliser

    @property
    def nsec(self) -> int:
        return int(self.sections.shape[0])

    @property
    def r_max(self) -> float:
        return float(np.max(np.abs(self.poles)))

    def ba(self) -> tuple[np.ndarray, np.ndarray]:
        """Direct-form coefficients (what scipy's `bessel(..., 'low')` returns);
        used only by tests to compare with the reference's design."""
        a = np.real(np.poly(self.poles))
        b = self.gain * np.array([math.comb(self.order, k) for k in range(self.order + 1)], dtype=np.float64)
        return b, a

    def impulse_tail(self, eps: float, nmax: int = 1 << 20) -> int:
        """Smallest n with sum_{m>=n} |h[m]| < eps (h = causal impulse response).
        This is the IIR warm-up halo length H: starting the recursion n samples
        early from a wrong state leaves an error below eps * max|x - median|."""
        cache = self.__dict__.setdefault("_tail_cache", {})
        key = (float(eps), int(nmax))
        if key not in cache:
            cache[key] = self._impulse_tail(eps, nmax)
        return cache[key]

    def _impulse_tail(self, eps: float, nmax: int) -> int:
        n = 4096
        while True:
            h = self._impulse(n)
            tail = np.cumsum(np.abs(h)[::-1])[::-1]
            # require that the computed window itself has converged
            if tail[-1] < eps * 1e-3 or n >= nmax:
                idx = np.nonzero(tail < eps)[0]
                return int(idx[0]) if idx.size else n
            n *= 4

    def _impulse(self, n: int) -> np.ndarray:
        x = np.zeros(n)
        x[0] = 1.0
        for (a1, a2), fo in zip(self.sections, self.first_order):
            v = np.empty(n)
            v1 = v2 = 0.0
            for i in range(n):  # n is a few thousand; host-only, once per design
                t = x[i] - a1 * v1 - a2 * v2
                v[i] = t
                v2, v1 = v1, t
            if fo:
                y = v.copy()
                y[1:] += v[:-1]
            else:
                y = v.copy()
                y[1:] += 2 * v[:-1]
                y[2:] += v[:-2]
            x = y
        return x * self.gain


_design_cache: dict[tuple[int, float], BesselDesign] = {}


def bessel_lowpass(order: int, wn: float) -> BesselDesign:
    """Digital low-pass Bessel of `order` with critical frequency `wn` (in units
    of Nyquist, exactly the `Wn = 2*cutoff/samplerate` of plot-trace.py:316)."""
    order = int(order)
    wn = float(wn)
    if not (0.0 < wn < 1.0):
        raise ValueError("Digital filter critical frequencies must be 0 < Wn < 1")
    key = (order, wn)
    if key in _design_cache:
        return _design_cache[key]
    pa = bessel_analog_poles(order)
    warped = 4.0 * math.tan(math.pi * wn / 2.0)
    pa = pa * warped
    pd = (4.0 + pa) / (4.0 - pa)
    # sections: complex pairs (imag > 0 member) sorted by ascending radius, real pole last
    cpx = [p for p in pd if p.imag > 1e-14]
    real = [p for p in pd if abs(p.imag) <= 1e-14]
    cpx.sort(key=lambda p: abs(p))
    secs, fo, sg = [], [], []
    for p in cpx:
        a1, a2 = -2.0 * p.real, p.real * p.real + p.imag * p.imag
        secs.append((a1, a2)); fo.append(False); sg.append((1.0 + a1 + a2) / 4.0)
    for p in real:
        a1 = -p.real
        secs.append((a1, 0.0)); fo.append(True); sg.append((1.0 + a1) / 2.0)
    if len(secs) > MAX_SECTIONS:
        raise ValueError("unsupported filter order")
    d = BesselDesign(order=order, wn=wn, poles=np.asarray(pd, dtype=np.complex128),
                     sections=np.asarray(secs, dtype=np.float64).reshape(-1, 2),
                     first_order=np.asarray(fo, dtype=bool),
                     gain=float(np.prod(sg)), section_gain=np.asarray(sg, dtype=np.float64))
    if d.r_max >= 1.0:
        raise ValueError("designed filter is unstable")
    _design_cache[key] = d
    return d
