"""TEST INFRASTRUCTURE — ctypes access to oracle/libct_oracle.so (the plain-C twin)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def build() -> str:
    path = os.path.join(_HERE, "libct_oracle.so")
    src = os.path.join(_HERE, "ct_oracle.c")
    if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "libct_oracle.so"], stdout=subprocess.DEVNULL)
    return path


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_detect.restype = C.c_int64
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def lfilter(b, a, x, z):
    b = np.ascontiguousarray(b, np.float64); a = np.ascontiguousarray(a, np.float64)
    x = np.ascontiguousarray(x, np.float64); z = np.array(z, np.float64)
    y = np.empty_like(x)
    lib().orc_lfilter(_p(b, C.c_double), _p(a, C.c_double), C.c_int(len(a)), _p(x, C.c_double),
                      C.c_int64(x.size), _p(z, C.c_double), _p(y, C.c_double))
    return y, z


def block_stats(y, block, bmin, bmax, c0, shift):
    y = np.ascontiguousarray(y, np.float32)
    nb = (y.size + block - 1) // block
    cnt = np.zeros(nb, np.int64); s1 = np.zeros(nb, np.int64); s2 = np.zeros(nb, np.int64)
    lib().orc_block_stats(_p(y, C.c_float), C.c_int64(y.size), C.c_int64(block), C.c_float(bmin),
                          C.c_float(bmax), C.c_float(c0), C.c_int(shift), _p(cnt, C.c_int64),
                          _p(s1, C.c_int64), _p(s2, C.c_int64))
    return cnt, s1, s2


def detect_events(y, block, sign, t_start, t_end, state_in=False, cap=None):
    y = np.ascontiguousarray(y, np.float32)
    sign = np.ascontiguousarray(sign, np.int32)
    t_start = np.ascontiguousarray(t_start, np.float32); t_end = np.ascontiguousarray(t_end, np.float32)
    cap = int(cap if cap is not None else max(16, y.size // 2))
    starts = np.zeros(cap, np.int64); ends = np.zeros(cap, np.int64)
    op = C.c_int64(-1)
    ne = lib().orc_detect(_p(y, C.c_float), C.c_int64(y.size), C.c_int64(block), _p(sign, C.c_int32),
                          _p(t_start, C.c_float), _p(t_end, C.c_float), C.c_int(int(state_in)),
                          _p(starts, C.c_int64), _p(ends, C.c_int64), C.c_int64(cap), C.byref(op))
    if ne > cap:
        raise RuntimeError("event capacity exceeded")
    return starts[:ne].copy(), ends[:ne].copy(), int(op.value)


def cusum_batch(samples, offsets, delta, h, max_levels=32):
    samples = np.ascontiguousarray(samples, np.float32)
    offsets = np.ascontiguousarray(offsets, np.int64)
    E = offsets.size - 1
    nlev = np.zeros(E, np.int32); edges = np.full((E, max_levels + 1), -1, np.int32)
    mean = np.zeros((E, max_levels)); std = np.zeros((E, max_levels)); ovf = np.zeros(E, np.uint8)
    lib().orc_cusum_batch(_p(samples, C.c_float), _p(offsets, C.c_int64), C.c_int64(E), C.c_float(delta),
                          C.c_float(h), C.c_int(max_levels), _p(nlev, C.c_int32), _p(edges, C.c_int32),
                          _p(mean, C.c_double), _p(std, C.c_double), _p(ovf, C.c_uint8))
    return nlev, edges, mean, std, ovf
