"""cusumtools_b200 — B200-native implementation of the raw-trace hot path of
shadowk29/cusumtools (load -> dequantise -> Bessel filtfilt -> threshold detection ->
CUSUM+ levels, plus Welch PSD).  Hand-written sm_100a CUDA behind a C ABI
(include/cusumtools_b200.h); this package is the Python mirror of the reference's entry
points.  There is no CPU fallback."""
__version__ = "0.1.0"
