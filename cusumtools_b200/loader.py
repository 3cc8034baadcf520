"""Trace loaders: the reference's file formats -> device tensors.

Mirrors `App.get_filenames` / `App.load_memmaps` / `App.get_file_index` /
`App.load_mapped_data` / `App.initialize_samplerate` (plot-trace.py:220-307,325-327), the
`.bin` slice of print_trace.py:32-39 / noise-fit.py:89-91 and the legacy `>i2` records of
legacy/minimal_psd.py:188-193.  Files are memory-mapped exactly like the reference; the
selected window goes to the GPU as raw bytes (pinned staging buffer) and is decoded there.
"""
from __future__ import annotations

import glob
import os
import time
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib, filters


def _settings_equal(a, b) -> bool:
    keys = ("SETUP_TIAgain", "SETUP_preADCgain", "SETUP_pAoffset", "SETUP_ADCVREF", "SETUP_ADCBITS")
    return all(float(np.squeeze(a[k])) == float(np.squeeze(b[k])) for k in keys)


def _to_device(arr: np.ndarray, device, dtype) -> torch.Tensor:
    """host numpy (possibly a memmap slice) -> pinned staging -> device, viewed as dtype."""
    flat = np.ascontiguousarray(arr).view(np.uint8).reshape(-1)
    host = torch.empty(flat.size, dtype=torch.uint8, pin_memory=torch.cuda.is_available())
    host.numpy()[:] = flat
    return host.to(device, non_blocking=True).view(dtype)


@dataclass
class Piece:
    codes: np.ndarray        # uint16 view into the memmap
    settings: dict
    file_index: int


class ChimeraSeries:
    """A series of Chimera `.log` files + sibling `.mat` settings."""

    def __init__(self, initialfile: str):
        import scipy.io as sio
        # get_filenames, plot-trace.py:301-307
        pattern = initialfile[:-19] + "*.log"
        files = glob.glob(pattern)
        stamps = [time.mktime(time.strptime(os.path.basename(f)[-19:-4], "%Y%m%d_%H%M%S")) for f in files]
        self.sorted_files = [f for (_, f) in sorted(zip(stamps, files), key=lambda p: p[0])]
        if not self.sorted_files:
            raise FileNotFoundError(f"no files match {pattern}")
        # load_memmaps, plot-trace.py:289-299
        ctype = np.dtype([("current", np.uint16)])
        self.maps = [np.memmap(str(f), dtype=ctype, mode="r")["current"] for f in self.sorted_files]
        self.settings = [sio.loadmat(f.replace(".log", ".mat")) for f in self.sorted_files]
        fsi = [0]
        for m in self.maps:
            fsi.append(fsi[-1] + len(m))
        self.total_samples = fsi[-1]
        self.file_start_index = np.array(fsi, dtype=np.int64)
        # initialize_samplerate, plot-trace.py:325-327
        self.samplerate = float(np.floor(np.squeeze(self.settings[0]["ADCSAMPLERATE"])))
        self.rate_mismatch = any(float(np.floor(np.squeeze(s["ADCSAMPLERATE"]))) != self.samplerate for s in self.settings)

    def get_file_index(self, samplenum: int) -> int:
        """plot-trace.py:220-227 (clamped instead of returning an out-of-range index)."""
        i = 0
        while i + 1 < len(self.file_start_index) - 1 and self.file_start_index[i + 1] < samplenum:
            i += 1
        return i

    def window(self, start_s: float | None, end_s: float | None) -> list[Piece]:
        """load_mapped_data's slicing (plot-trace.py:230-269): [start_s, end_s) -> file pieces."""
        fs = self.samplerate
        start_index = int(float(start_s) * fs) if start_s is not None else 0
        start_f = self.get_file_index(start_index)
        if end_s is not None:
            end_index = min(int(float(end_s) * fs), self.total_samples)
        else:
            end_index = min(start_index + len(self.maps[start_f]), self.total_samples)
        end_f = self.get_file_index(end_index)
        fsi = self.file_start_index
        if start_f == end_f:
            return [Piece(self.maps[start_f][start_index - fsi[start_f]:end_index - fsi[start_f]], self.settings[start_f], start_f)]
        pieces = [Piece(self.maps[start_f][start_index - fsi[start_f]:], self.settings[start_f], start_f)]
        for i in range(start_f + 1, end_f):
            pieces.append(Piece(self.maps[i], self.settings[i], i))
        pieces.append(Piece(self.maps[end_f][:end_index - fsi[end_f]], self.settings[end_f], end_f))
        return pieces

    def host_codes(self, start_s=None, end_s=None):
        """Raw codes of the window gathered into ONE pinned host tensor (what
        pipeline.StreamingAnalyzer.run_from_host streams to the GPU) plus the settings; requires one
        gain for the whole window (the common case).  Returns (uint16 CPU tensor, settings)."""
        pieces = self.window(start_s, end_s)
        if not all(_settings_equal(p.settings, pieces[0].settings) for p in pieces):
            raise ValueError("files in the window have different gain settings: use load_pA()")
        n = sum(len(p.codes) for p in pieces)
        host = torch.empty(n, dtype=torch.uint16, pin_memory=torch.cuda.is_available())
        hv = host.numpy()
        o = 0
        for p in pieces:
            hv[o:o + len(p.codes)] = p.codes
            o += len(p.codes)
        return host, pieces[0].settings

    def reader(self, start_s=None, end_s=None) -> "WindowReader":
        """The window as a random-access source of raw codes for the streamed analysis
        (pipeline.StreamingAnalyzer.run_from_file): nothing is read until a range is asked for."""
        pieces = self.window(start_s, end_s)
        if not all(_settings_equal(p.settings, pieces[0].settings) for p in pieces):
            raise ValueError("files in the window have different gain settings: use load_pA()")
        fs = self.samplerate
        start_index = int(float(start_s) * fs) if start_s is not None else 0
        spans, o = [], 0
        for p in pieces:                         # (window offset, file, first sample inside the file, samples)
            first = start_index - int(self.file_start_index[p.file_index]) if p is pieces[0] else 0
            spans.append((o, self.sorted_files[p.file_index], first, len(p.codes)))
            o += len(p.codes)
        return WindowReader(spans, o, pieces[0].settings)

    def load_codes(self, start_s=None, end_s=None, device="cuda"):
        """`host_codes` copied to the device.  Returns (uint16 tensor, settings)."""
        host, settings = self.host_codes(start_s, end_s)
        return host.to(device, non_blocking=True), settings

    def load_pA(self, start_s=None, end_s=None, device="cuda") -> torch.Tensor:
        """load_mapped_data itself: every piece scaled with its own file's settings
        (plot-trace.py:252-269), float32 pA on the device."""
        pieces = self.window(start_s, end_s)
        n = sum(len(p.codes) for p in pieces)
        out = torch.empty(n, dtype=torch.float32, device=device)
        L = _lib.lib()
        o = 0
        for p in pieces:
            raw = _to_device(p.codes, device, torch.uint16)
            alpha, beta = filters.chimera_affine(p.settings)
            rc = L.ct_dequant_u16(raw.data_ptr(), raw.numel(), filters.chimera_bitmask(p.settings), alpha, beta,
                                  out[o:].data_ptr(), filters._stream_ptr(out))
            _lib.check(rc, "ct_dequant_u16")
            o += raw.numel()
        return out


class WindowReader:
    """Raw uint16 codes of a window of a `.log` series (plot-trace.py:230-269 slicing), readable range by range
    straight into caller memory (pinned slabs): `os.preadv` copies from the page cache / disk into the destination
    in one step and releases the GIL, so several threads can fill disjoint parts of a slab at once."""

    def __init__(self, spans, n: int, settings):
        self.spans, self.n, self.settings = spans, int(n), settings
        self._fds: dict = {}

    def _fd(self, path: str) -> int:
        fd = self._fds.get(path)
        if fd is None:
            fd = self._fds[path] = os.open(path, os.O_RDONLY)
        return fd

    def read_into(self, dst: np.ndarray, a: int, b: int) -> None:
        """dst[0 : b - a] = window samples [a, b) (dst: contiguous uint16/int16 array of at least b - a entries)."""
        mv = memoryview(dst.view(np.uint8).reshape(-1))
        for off, path, first, cnt in self.spans:
            lo, hi = max(a, off), min(b, off + cnt)
            if hi <= lo:
                continue
            pos, want = 2 * (first + lo - off), 2 * (hi - lo)
            out = mv[2 * (lo - a):2 * (lo - a) + want]
            got = 0
            while got < want:
                r = os.preadv(self._fd(path), [out[got:]], pos + got)
                if r <= 0:
                    raise IOError(f"short read from {path}")
                got += r

    def close(self) -> None:
        for fd in self._fds.values():
            os.close(fd)
        self._fds = {}

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def load_chimera(path: str, start_s=None, end_s=None, device="cuda"):
    """(codes uint16 CUDA tensor, settings, samplerate) for a window of a `.log` series."""
    s = ChimeraSeries(path)
    raw, settings = s.load_codes(start_s, end_s, device)
    return raw, settings, s.samplerate


def load_bin(path: str, start_s: float | None = None, length_s: float | None = None,
             samplingfreq: float = 4166666.0, device="cuda") -> torch.Tensor:
    """`.bin` records (>f8 curr_pA, >f8 volt_mV): current[int(start_s*fs):int((start_s+length_s)*fs)]
    (print_trace.py:33-39) as float32 pA on the device; the whole file when start_s is None."""
    rec = np.memmap(path, dtype=np.dtype([("curr_pA", ">f8"), ("volt_mV", ">f8")]), mode="r")
    if start_s is not None:
        rec = rec[int(start_s * samplingfreq):int((start_s + length_s) * samplingfreq)]
    n = len(rec)
    out = torch.empty(n, dtype=torch.float32, device=device)
    if n:
        raw = _to_device(rec, device, torch.uint8)
        rc = _lib.lib().ct_bin_be_f64_to_f32(raw.data_ptr(), n, out.data_ptr(), filters._stream_ptr(out))
        _lib.check(rc, "ct_bin_be_f64_to_f32")
    return out


def load_legacy_i2(path: str, start: int, end: int, savegain: float, device="cuda") -> torch.Tensor:
    """legacy/minimal_psd.py:188-193: savegain * current[start:end] from (>i2, >i2) records."""
    rec = np.memmap(path, dtype=np.dtype([("current", ">i2"), ("voltage", ">i2")]), mode="r")[start:end]
    n = len(rec)
    out = torch.empty(n, dtype=torch.float32, device=device)
    if n:
        raw = _to_device(rec, device, torch.uint8)
        rc = _lib.lib().ct_i2be_to_f32(raw.data_ptr(), n, float(savegain), out.data_ptr(), filters._stream_ptr(out))
        _lib.check(rc, "ct_i2be_to_f32")
    return out
