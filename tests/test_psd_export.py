"""PSD export formats the reference's consumers read: the `f,Pxx,rms`
CSV of App.export_psd (plot-trace.py:204-207) and the 4-column `.psd` TSV that
legacy/psdfit.py:27-30 opens with pd.read_csv(sep='\\t', names=['f','S','integral','norm'])."""
import numpy as np
import pandas as pd

from cusumtools_b200 import psd


def test_export_formats(tmp_path):
    f = np.arange(513) * (4166666.0 / 1024)
    P = 1e-3 / (1 + np.arange(513))
    rms = psd.integrate_noise(f, P)                                   # plot-trace.py:309-311
    p = tmp_path / "psd.csv"
    psd.export_psd(str(p), f, P, rms)
    back = np.loadtxt(p, delimiter=",")
    assert back.shape == (513, 3) and np.array_equal(back, np.c_[f, P, rms])
    q = tmp_path / "trace.psd"
    psd.export_psd_tsv(str(q), f, P, current=5000.0, bandwidth=100000.0)
    tab = pd.read_csv(q, sep="\t", names=["f", "S", "integral", "norm"])       # legacy/psdfit.py:27
    assert np.allclose(tab["f"].values, f, rtol=1e-15) and np.allclose(tab["S"].values, P, rtol=1e-15)
    assert np.allclose(np.sqrt(tab["integral"].values), rms, rtol=1e-14)
    assert np.allclose(tab["norm"].values, P / 5000.0 ** 2 * 1e5, rtol=1e-14)  # plot-trace.py:445-447
    fx = tab["f"].values[1:100]
    assert np.isfinite(np.log10(tab["norm"].values[1:100])).all() and fx[1] - fx[0] > 0   # what psdfit.py:28-31 uses
