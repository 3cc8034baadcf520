"""Headless mirrors of the reference's entry-point classes, method for method, on top of the CUDA path.

The reference has no plugin interface; its operator boundary is the method set of its Tk `App` classes
and of `SpectrumSample`.  The classes below keep those names, argument meanings, attribute
names and the error behaviour (status strings in `wildcard`, the same exceptions where the reference lets
one escape), without Tk and matplotlib: an "entry" is any object with `get()`, what the reference plots is
left in attributes.  Code written against `plot-trace.py`'s `App` (`app.cutoff_entry`, `app.update_psd()`,
`app.f / app.Pxx / app.rms`, `app.filtered_data`) runs unchanged.

  App              plot-trace.py:55-527        Chimera `.log` series -> trace / PSD / overlays
  LegacyPsdApp     legacy/minimal_psd.py:52-293  (`>i2`, `>i2`) records, savegain, 2^18-point Welch
  BesselStepApp    legacy/bessel-filter.py:29-137  step response with the edge + odd-extension boundary mode
  SpectrumSample   noise-fit.py:84-113         `.bin` trace -> cropped, normalised |I| spectrum
  print_trace      print_trace.py:24-41        slice of a `.bin` trace -> `%.18e` CSV

Arrays that hold samples (`data`, `filtered_data`, `plot_data`, `downsampled_data`) are float32 CUDA tensors
(the reference's are float64 numpy; tolerances in DESIGN.md 2); spectra are float64 numpy like scipy's.
There is no CPU fallback: every method that computes needs the CUDA library."""
from __future__ import annotations

import re

import numpy as np
import torch

from . import _lib, filters, loader, psd


class Entry:
    """Stand-in for tk.Entry / tk.StringVar / tk.IntVar: `get`, `set`, `insert`, `delete`."""

    def __init__(self, value=""):
        self._v = value

    def get(self):
        return self._v

    def set(self, value) -> None:
        self._v = value

    def insert(self, index, text) -> None:
        s = str(self._v)
        i = len(s) if index in ("end", "END") else int(index)
        self._v = s[:i] + str(text) + s[i:]

    def delete(self, first, last=None) -> None:
        s = str(self._v)
        a = int(first)
        b = len(s) if last in ("end", "END") else (a + 1 if last is None else int(last))
        self._v = s[:a] + s[b:]


def _to_numpy(t) -> np.ndarray:
    return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


class _PsdMixin:
    def integrate_noise(self, f, Pxx):
        """plot-trace.py:309-311, legacy/minimal_psd.py:195-197."""
        return psd.integrate_noise(f, Pxx)

    def export_psd(self, data_path: str) -> None:
        """plot-trace.py:206-211: rows `f,Pxx,rms`; 'Plot the PSD first' when there is none."""
        try:
            psd.export_psd(data_path, self.f, self.Pxx, self.rms)
        except AttributeError:
            self.wildcard.set("Plot the PSD first")

    def export_trace(self, data_path: str) -> None:
        """plot-trace.py:213-218."""
        try:
            np.savetxt(data_path, _to_numpy(self.plot_data), delimiter=",")
        except AttributeError:
            self.wildcard.set("Plot the trace first")

    def _filter_requested(self) -> bool:
        # `self.order_entry != ''` compares the widget, not its text (plot-trace.py:336): always true
        return self.cutoff_entry.get() != ""


class App(_PsdMixin):
    """`plot-trace.py`'s App without the GUI (plot-trace.py:55-527)."""

    def __init__(self, parent, file_path: str, device="cuda"):
        self.device = device
        self.events_flag = False
        self.baseline_flag = False
        self.overlay_flag = False
        self.file_path = file_path
        self.start_entry = Entry("0")
        self.end_entry = Entry("10")
        self.psd_length_entry = Entry("")
        self.cutoff_entry = Entry("900000")
        self.order_entry = Entry("8")
        self.downsample_entry = Entry("")
        self.normalize = Entry(0)
        self.wildcard = Entry("")
        self._codes = self._code_settings = self._data_src = None
        self.update_data()

    # ------------------------------------------------------------------ loaders
    def update_data(self) -> None:
        """plot-trace.py:523-526."""
        self.get_filenames(self.file_path)
        self.load_memmaps()
        self.initialize_samplerate()

    def get_filenames(self, initialfile: str) -> None:
        """plot-trace.py:301-307.  No matching file: FileNotFoundError here (the reference reports 'Found 0
        files' and fails with an IndexError in initialize_samplerate)."""
        self.series = loader.ChimeraSeries(initialfile)
        self.sorted_files = self.series.sorted_files
        self.wildcard.set("Found {0} files matching {1}".format(len(self.sorted_files), initialfile[:-19] + "*.log"))

    def load_memmaps(self) -> None:
        """plot-trace.py:289-299."""
        s = self.series
        self.maps, self.settings = s.maps, s.settings
        self.file_start_index, self.total_samples = s.file_start_index, s.total_samples

    def initialize_samplerate(self) -> None:
        """plot-trace.py:325-327."""
        self.samplerate = np.floor(np.squeeze(self.settings[0]["ADCSAMPLERATE"]))

    def get_file_index(self, samplenum: int) -> int:
        """plot-trace.py:220-227, clamped to the last file (the reference returns an out-of-range index)."""
        return self.series.get_file_index(samplenum)

    def scale_raw_data(self, tempdata, settings) -> torch.Tensor:
        """plot-trace.py:272-287: raw uint16 codes (numpy or tensor) -> pA, float32 on the device
        (`ct_dequant_u16`: the float64 affine of the reference's sequence, rounded once)."""
        samplerate = np.floor(np.squeeze(settings["ADCSAMPLERATE"]))
        if samplerate != self.samplerate:
            self.wildcard.set("One of your files does not match the global sampling rate!")
        if isinstance(tempdata, torch.Tensor):
            raw = tempdata.to(self.device)
            if raw.dtype == torch.int16:
                raw = raw.view(torch.uint16)
        else:
            raw = loader._to_device(np.ascontiguousarray(np.asarray(tempdata).astype(np.uint16)), self.device, torch.uint16)
        out = torch.empty(raw.numel(), dtype=torch.float32, device=raw.device)
        if raw.numel():
            alpha, beta = filters.chimera_affine(settings)
            with torch.cuda.device(raw.device):
                rc = _lib.lib().ct_dequant_u16(raw.data_ptr(), raw.numel(), filters.chimera_bitmask(settings), alpha, beta,
                                               out.data_ptr(), filters._stream_ptr(out))
            _lib.check(rc, "ct_dequant_u16")
        return out

    def load_mapped_data(self) -> None:
        """plot-trace.py:230-270: the window [start_entry, end_entry) in seconds, every file piece scaled with
        its own settings.  The raw codes of a single-gain window stay on the device so that `filter_data` can
        run the fused dequantise + filter kernels on them."""
        if self.start_entry.get() != "":
            self.start_time = float(self.start_entry.get())
            start_s = self.start_time
        else:
            self.start_time = 0
            start_s = None
        if self.end_entry.get() != "":
            self.end_time = float(self.end_entry.get())
            end_s = self.end_time
        else:
            end_s = None
        pieces = self.series.window(start_s, end_s)
        self._codes = self._code_settings = None
        if all(loader._settings_equal(p.settings, pieces[0].settings) for p in pieces):
            self._codes, self._code_settings = self.series.load_codes(start_s, end_s, self.device)
            data = self.scale_raw_data(self._codes, self._code_settings)
        else:
            data = torch.cat([self.scale_raw_data(p.codes, p.settings) for p in pieces])
        self.data = self._data_src = data

    # ------------------------------------------------------------------ filter
    def filter_data(self) -> None:
        """plot-trace.py:313-320: np.pad(mode='median', 1000) + filtfilt(bessel(order, 2 fc / fs), padtype=None)."""
        cutoff = float(self.cutoff_entry.get())
        order = int(self.order_entry.get())
        if self._codes is not None and self.data is self._data_src:
            self.filtered_data = filters.dequant_filtfilt(self._codes, self._code_settings, cutoff, order,
                                                          samplerate=float(self.samplerate))
        else:
            self.filtered_data = filters.bessel_filtfilt(self._as_device(self.data), float(self.samplerate), cutoff, order)

    def _as_device(self, x) -> torch.Tensor:
        if isinstance(x, torch.Tensor):
            return x.to(self.device, torch.float32)
        return torch.from_numpy(np.ascontiguousarray(np.asarray(x), dtype=np.float32)).to(self.device)

    def downsample_data(self) -> None:
        """plot-trace.py:322-323 (with the entry converted to a number; the reference divides by the string)."""
        self.downsampled_data = self.filtered_data[::int(self.samplerate / float(self.downsample_entry.get()))]

    # ------------------------------------------------------------------ overlays
    def overlay_cusum(self, analysis_dir: str) -> None:
        """plot-trace.py:172-203: rate.csv, baseline.csv and the detector settings of an analysis directory."""
        import pandas as pd
        baseline_path = analysis_dir + "/baseline.csv"
        ratefile_path = analysis_dir + "/rate.csv"
        config_path = analysis_dir + "/summary.txt"
        self.events_flag = True
        self.baseline_flag = True
        self.overlay_flag = True
        try:
            self.ratefile = pd.read_csv(ratefile_path, encoding="utf-8")
        except IOError:
            self.overlay_flag = False
            self.wildcard.set("rate.csv not found in given directory")
        try:
            self.baseline_file = pd.read_csv(baseline_path, encoding="utf-8")
        except IOError:
            self.overlay_flag = False
            self.wildcard.set("baseline.csv not found in given directory")
        # summary.txt: `key=value` lines matched by SUBSTRING like the reference does (so `intra_threshold` must not be
        # taken for `threshold`): key fragment, attribute, type, whether an `intra` line is skipped
        wanted = (("threshold", "threshold", float, True), ("hysteresis", "hysteresis", float, True),
                  ("cutoff", "config_cutoff", int, False), ("poles", "config_order", int, False))
        with open(config_path, "r") as config:
            for line in config:
                for fragment, attr, kind, skip_intra in wanted:
                    if fragment in line and not (skip_intra and "intra" in line):
                        setattr(self, attr, kind(re.split("=|\n", line)[1]))

    def _event_spans(self):
        """The axvspan lists of plot-trace.py:350-369 in microseconds: (good, bad) lists of (start, end)."""
        db, t0, t1 = self.ratefile, self.start_time, self.end_time
        ok = db["type"].isin((0, 1))
        s_in = (db["start_time_s"] >= t0) & (db["start_time_s"] < t1)
        e_in = (db["end_time_s"] >= t0) & (db["end_time_s"] < t1)
        out = []
        for sel in (ok, db["type"] > 1):
            start = np.atleast_1d(db.loc[s_in & sel, "start_time_s"].to_numpy(dtype=float) * 1e6)
            end = np.atleast_1d(db.loc[e_in & sel, "end_time_s"].to_numpy(dtype=float) * 1e6)
            if len(start) > 0 and len(end) > 0 and start[0] > end[0]:
                start = start[1:]
            out.append(list(zip(start, end)))
        return out[0], out[1]

    def _baseline_lines(self):
        """plot-trace.py:379-414: per baseline block (xmin_us, xmax_us, baseline, hysteresis line, threshold line)."""
        db, start_time, end_time = self.baseline_file, self.start_time, self.end_time
        times = np.sort(np.atleast_1d(np.squeeze(db[["time_s"]].values)))
        start_block = times[0]
        for t in times:
            if t <= start_time and t >= start_block:
                start_block = t
        sel = db[(db["time_s"] >= start_block) & (db["time_s"] < end_time)]
        times, means, stdevs = sel["time_s"].values, sel["baseline_pA"].values, sel["stdev_pA"].values
        lines = []
        for i in range(len(means)):
            xmin = start_time if i == 0 else times[i]
            xmax = end_time if i + 1 == len(means) else times[i + 1]
            sign = np.sign(means[i])
            lines.append((xmin * 1e6, xmax * 1e6, means[i], means[i] - sign * (self.threshold - self.hysteresis) * stdevs[i],
                          means[i] - sign * self.threshold * stdevs[i]))
        return lines

    # ------------------------------------------------------------------ the two buttons
    def update_trace(self) -> None:
        """plot-trace.py:330-415 without the drawing: `plot_data` (+ `plot_samplerate`, `plot_time_us()`), the
        event spans (`good_spans`, `bad_spans`) and the baseline / threshold lines (`baseline_lines`)."""
        self.load_mapped_data()
        self.filtered_data = self.data
        self.plot_data = self.filtered_data
        self.plot_samplerate = self.samplerate
        if self._filter_requested():
            self.filter_data()
            self.plot_data = self.filtered_data
        if self.downsample_entry.get() != "":
            self.downsample_data()
            self.plot_data = self.downsampled_data
            self.plot_samplerate = float(self.downsample_entry.get())
        self.good_spans, self.bad_spans, self.baseline_lines = [], [], []
        if self.events_flag and self.overlay_flag:
            self.good_spans, self.bad_spans = self._event_spans()
        if self.baseline_flag:
            if self.config_cutoff != int(self.cutoff_entry.get()) or self.config_order != int(self.order_entry.get()):
                self.wildcard.set("Filter settings in config file do not match plotting filter settings, overlay will be inaccurate")
            self.baseline_lines = self._baseline_lines()

    def plot_time_us(self) -> np.ndarray:
        """The x axis of the trace plot (plot-trace.py:372,378)."""
        n, fs = len(self.plot_data), float(self.plot_samplerate)
        return (np.linspace(1.0 / fs, n / fs, n) + self.start_time) * 1e6

    def update_psd(self) -> None:
        """plot-trace.py:418-509 without the drawing: `f`, `Pxx`, `rms`, `current`, the plot limits in
        `psd_limits` = (minf, maxf, minP, maxP) and the fit inputs `fnorm`, `Pxx_norm` (f < 10 kHz)."""
        self.load_mapped_data()
        self.filtered_data = self.data
        self.plot_data = self.filtered_data
        plot_samplerate = float(self.samplerate)
        bandwidth = 1.0e6
        cutoff = None
        if self._filter_requested():
            self.filter_data()
            self.plot_data = self.filtered_data
            cutoff = float(self.cutoff_entry.get())
            maxf = 2 * cutoff
            bandwidth = maxf / 2.0
        else:
            maxf = 2e6
        psd_length_s = float(self.psd_length_entry.get()) if self.psd_length_entry.get() != "" else None
        normalize = bool(self.normalize.get())
        f, Pxx, rms, current = psd.update_psd(self._as_device(self.filtered_data), plot_samplerate,
                                              psd_length_s=psd_length_s, normalize=normalize, cutoff=cutoff)
        self.f, self.Pxx, self.rms, self.current = f, Pxx, rms, current
        BW_index = np.searchsorted(f, maxf / 2)
        logPxx = np.log10(Pxx[1:BW_index])
        self.psd_limits = (1, maxf, 10 ** np.floor(np.amin(logPxx)), 10 ** np.ceil(np.amax(logPxx)))
        N = len(f[f < 10000])
        self.fnorm = f[1:N]
        self.Pxx_norm = Pxx[1:N] if normalize else Pxx[1:N] * bandwidth / current ** 2

    # ------------------------------------------------------------------ fit helpers (plot-trace.py:511-521)
    def fitfunc(self, f, f0, alpha, fstar, offset):
        return np.log10((f0 / f) ** alpha + alpha * (f0 / fstar) ** (1 + alpha) * (f / f0) + offset)

    def corrected_L(self, f, Pxx, f0, alpha, fstar, offset, df, B):
        integrand = Pxx - alpha * (f0 / fstar) ** (1 + alpha) * (f / f0) - offset
        return np.sqrt(np.sum(integrand) * df / B)

    def old_L(self, Pxx, df, B):
        return np.sqrt(np.sum(Pxx) * df / B)


class LegacyPsdApp(_PsdMixin):
    """`legacy/minimal_psd.py`'s App without the GUI (legacy/minimal_psd.py:52-293): one file of big-endian
    (`>i2` current, `>i2` voltage) records, scaled by `savegain_entry`, sample rate from `samplerate_entry`."""

    def __init__(self, parent, file_path: str, device="cuda"):
        self.device = device
        self.file_path = file_path
        self.samplerate_entry = Entry("")
        self.savegain_entry = Entry("")
        self.start_entry = Entry("0")
        self.end_entry = Entry("10")
        self.cutoff_entry = Entry("")
        self.order_entry = Entry("")
        self.normalize = Entry(0)
        self.wildcard = Entry("")
        self.load_memmap()

    def load_memmap(self) -> None:
        """legacy/minimal_psd.py:191-193."""
        columntypes = np.dtype([("current", ">i2"), ("voltage", ">i2")])
        self.map = np.memmap(self.file_path, dtype=columntypes, mode="r")["current"]

    def initialize_samplerate(self) -> None:
        """legacy/minimal_psd.py:208-209."""
        self.samplerate = float(self.samplerate_entry.get())

    def load_mapped_data(self) -> None:
        """legacy/minimal_psd.py:171-189: savegain * map[start:end] (`ct_i2be_to_f32`).  With an empty end
        entry the reference fails on the unbound `end_index`; so does this (UnboundLocalError)."""
        self.total_samples = len(self.map)
        self.samplerate = int(self.samplerate_entry.get())
        if self.start_entry.get() != "":
            self.start_time = float(self.start_entry.get())
            start_index = int(float(self.start_entry.get()) * self.samplerate)
        else:
            self.start_time = 0
            start_index = 0
        if self.end_entry.get() != "":
            self.end_time = float(self.end_entry.get())
            end_index = int(float(self.end_entry.get()) * self.samplerate)
            if end_index > self.total_samples:
                end_index = self.total_samples
        self.data = loader.load_legacy_i2(self.file_path, start_index, end_index, float(self.savegain_entry.get()),
                                          device=self.device)

    def filter_data(self) -> None:
        """legacy/minimal_psd.py:199-206 (the same call sequence as plot-trace.py:313-320)."""
        self.filtered_data = filters.bessel_filtfilt(self.data, float(self.samplerate), float(self.cutoff_entry.get()),
                                                     int(self.order_entry.get()))

    def update_trace(self) -> None:
        """legacy/minimal_psd.py:212-234 without the drawing."""
        self.initialize_samplerate()
        self.load_mapped_data()
        self.filtered_data = self.data
        self.plot_data = self.filtered_data
        if self._filter_requested():
            self.filter_data()
            self.plot_data = self.filtered_data

    def update_psd(self) -> None:
        """legacy/minimal_psd.py:236-262: nperseg = min(2^18, len); normalisation by current^2 and maxf / 2."""
        self.initialize_samplerate()
        self.load_mapped_data()
        self.filtered_data = self.data
        self.plot_data = self.filtered_data
        plot_samplerate = float(self.samplerate)
        if self._filter_requested():
            self.filter_data()
            self.plot_data = self.filtered_data
            maxf = 2 * float(self.cutoff_entry.get())
        else:
            maxf = 2 * float(self.samplerate_entry.get())
        n = self.filtered_data.numel()
        length = int(np.minimum(2 ** 18, n))
        end_index = int(np.floor(n / length) * length)
        current = float(self.filtered_data[:end_index].to(torch.float64).mean().item())
        f, Pxx = psd.welch(self.filtered_data, plot_samplerate, length)
        self.rms = self.integrate_noise(f, Pxx)
        if self.normalize.get():
            Pxx = Pxx / current ** 2 * (maxf / 2.0)
        self.f, self.Pxx, self.current = f, Pxx, current


class BesselStepApp:
    """`legacy/bessel-filter.py`'s App without the GUI (legacy/bessel-filter.py:29-137): unit step through the
    Bessel filter with the edge pad + scipy-default odd extension (the a8' boundary mode).  `fc_entry`, `fs_entry`
    in kHz."""

    def __init__(self, parent=None, device="cuda"):
        self.device = device
        self.fc_entry = Entry("")
        self.fs_entry = Entry("")
        self.poles = Entry("")

    def generate_step(self, length: int) -> None:
        """legacy/bessel-filter.py:133-136 (`range(length/2, length)` with Python 2's integer division)."""
        self.perfect_data = torch.zeros(int(length), dtype=torch.float32, device=self.device)
        self.perfect_data[int(length) // 2:] = 1

    def filter_data(self) -> None:
        """legacy/bessel-filter.py:124-131."""
        fc = 1000 * float(self.fc_entry.get())
        fs = 1000 * float(self.fs_entry.get())
        self.filtered_data = filters.bessel_filtfilt_odd(self.perfect_data, fs, fc, int(self.poles.get()))

    def update_filter(self) -> None:
        """legacy/bessel-filter.py:100-105 (the step and its response; the curve fit and the plot are the GUI's)."""
        self.fc = 1000 * float(self.fc_entry.get())
        self.fs = 1000 * float(self.fs_entry.get())
        self.generate_step(int(5 * self.fs / self.fc))
        self.filter_data()


class SpectrumSample:
    """noise-fit.py:84-100: |current| spectrum of a `.bin` trace, cropped to f <= cutoff and scaled by
    cutoff / I^2 and by f.  `fit_spectrum` (noise-fit.py:102-110) is the reference's own host-side curve fit
    on those ~100 bins."""

    def __init__(self, tracefile: str, samplerate, psdlength, cutoff, device="cuda"):
        self.thermal = 1.0e-3
        self.pink = 1
        self.brown = 1.0e-3
        self.raw = loader.load_bin(tracefile, device=device)
        self.f, self.Pxx, self.current = psd.spectrum_sample(self.raw, float(samplerate), psdlength, cutoff)

    def fit_spectrum(self) -> None:
        from scipy.optimize import curve_fit           # a 3-parameter host fit, outside the hot path
        self.p0 = [self.thermal, self.pink, self.brown]
        with np.errstate(invalid="ignore"):            # the optimiser probes negative arguments of the log
            popt, _ = curve_fit(psd_fit, self.f, np.log10(self.Pxx), self.p0,
                                sigma=np.sqrt(np.arange(1, len(self.f) + 1) + np.sqrt(3) / 3), maxfev=100000)
        self.thermal, self.pink, self.brown = popt


def psd_fit(f, thermal, pink, brown):
    """noise-fit.py:8-10."""
    return np.log10(thermal * f + pink + brown / f)


def print_trace(file_path_string: str, start_s: float = 3270.207934, length_s: float = 0.264425,
                samplingfreq: float = 4166666, device="cuda") -> str:
    """print_trace.py:24-41: current[int(start_s fs) : int((start_s + length_s) fs)] of a `.bin` trace written as
    one `%.18e` value per line to `<name>_<start_s>_<length_s>_current.csv`; returns that path."""
    index = file_path_string.find(".bin")
    outname = file_path_string[:index] + "_" + str(start_s) + "_" + str(length_s) + "_current.csv"
    current = loader.load_bin(file_path_string, start_s, length_s, samplingfreq, device=device)
    np.savetxt(outname, current.cpu().numpy().astype(np.float64))
    return outname
