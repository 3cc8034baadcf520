"""The filter pair as the step runs it (forward pass with the exact-median window tally riding on it; backward pass with
block sums and chunk extrema), a few times,
for ncu:  ncu --set full -k regex:ct_filter -s 4 -c 2 python scripts/prof_filter.py [n] [cutoff]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from cusumtools_b200 import _lib, detect, filters, synth
from cusumtools_b200.design import bessel_lowpass

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 30
cutoff = float(sys.argv[2]) if len(sys.argv) > 2 else 1e5
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
S = synth.CHIMERA_SETTINGS
L = _lib.lib()
raw = synth.device_trace(n, "cuda", seed=1234)
out = torch.empty(n, dtype=torch.float32, device="cuda")
mask = filters.chimera_bitmask(S)
alpha, _ = filters.chimera_affine(S)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
d = bessel_lowpass(8, 2 * cutoff / synth.FS)
coef = filters.make_coef(d)
H = filters.warmup_samples(d)
wsb = int(L.ct_filtfilt_workspace_bytes(n, 1000, H))
ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
mm = torch.empty(2 * int(L.ct_filter_summary_count(n, 1000, H)), dtype=torch.float32, device="cuda")
bl = detect.new_baseline(n, 1 << 20, 4700.0, 5300.0, "cuda")
stats = detect.stats_args(bl, origin=0)
est, offset = 40900.0, 5000.0
cnt9 = torch.zeros(9, dtype=torch.int64, device="cuda")
for _ in range(reps):
    rc = L.ct_filter_forward_u16(raw.data_ptr(), n, 1000, est, mask, 0.0, C.byref(coef), H, 0, 0, int(est) - 12, 4, 0, n,
                                 cnt9.data_ptr(), 0, 0, ws.data_ptr(), wsb, st)
    assert rc == 0, L.ct_last_error()
    rc = L.ct_filter_backward(n, 1000, est, float(alpha), offset, C.byref(coef), H, 0, out.data_ptr(), ws.data_ptr(), wsb,
                              C.byref(stats), mm.data_ptr(), st)
    assert rc == 0, L.ct_last_error()
torch.cuda.synchronize()
print("done", float(out[12345].item()))
