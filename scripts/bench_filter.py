"""Kernel-only timing of ct_filtfilt_u16 (coefficients prepared once), for tuning."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cusumtools_b200 import _lib, filters, synth
from cusumtools_b200.design import bessel_lowpass
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 30
subs = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [4096]
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
raw = synth.device_trace(n, "cuda", seed=1)
out = torch.empty(n, dtype=torch.float32, device="cuda")
d = bessel_lowpass(8, 2 * 1e5 / synth.FS)
coef = filters.make_coef(d)
H = filters.halo_samples(d)
L = _lib.lib()
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
Hraw = d.impulse_tail(filters.DEFAULT_HALO_EPS)
wsb = L.ct_filtfilt_workspace_bytes(n, 1000, Hraw)
ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
for S in subs:
    def run():
        rc = L.ct_filtfilt_u16(raw.data_ptr(), n, 1000, 40900.0, 0xFFFC, 2.3385, 5000.0, C.byref(coef), S, H if S else Hraw, 0, out.data_ptr(), ws.data_ptr(), wsb, None, st)
        assert rc == 0, L.ct_last_error()
    run(); run(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"S={S} H={H} n={n}: {ms:.3f} ms  {n/ms/1e6:.1f} Gsamples/s  {6*n/ms/1e6:.0f} GB/s  frac {6*n/ms/1e6/6559.7:.3f}")
