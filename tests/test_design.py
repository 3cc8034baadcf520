"""Host-side Bessel design vs scipy.signal.bessel (the call at plot-trace.py:317)."""
import numpy as np
import pytest
from scipy.signal import bessel

from cusumtools_b200.design import bessel_lowpass


@pytest.mark.parametrize("order", list(range(1, 11)))
@pytest.mark.parametrize("wn", [0.0144, 0.048, 0.12, 0.432])
def test_design_matches_scipy(order, wn):
    d = bessel_lowpass(order, wn)
    b, a = bessel(order, wn, "low")
    b2, a2 = d.ba()
    assert np.allclose(a2, a, rtol=5e-11, atol=0)
    assert np.allclose(b2, b, rtol=5e-11, atol=0)
    assert d.nsec == (order + 1) // 2
    assert d.r_max < 1


def test_numerator_is_binomial():
    # B(z) = b0 (1 + z^-1)^8 exactly
    b, _ = bessel(8, 2 * 1e5 / 4166666.0, "low")
    assert np.allclose(b / b[0], [1, 8, 28, 56, 70, 56, 28, 8, 1], rtol=1e-9)


def test_halo_grows_as_cutoff_falls():
    h = [bessel_lowpass(8, 2 * fc / 4166666.0).impulse_tail(1e-7) for fc in (9e5, 2.5e5, 1e5, 3e4)]
    assert h == sorted(h) and 150 < h[2] < 400


def test_bad_arguments():
    with pytest.raises(ValueError):
        bessel_lowpass(8, 1.2)
    with pytest.raises(ValueError):
        bessel_lowpass(0, 0.1)
    with pytest.raises(ValueError):
        bessel_lowpass(11, 0.1)
