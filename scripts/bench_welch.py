"""Kernel-only timing of the Welch PSD (config C4 shape: 2^20-point segments)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cusumtools_b200 import psd, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 28
L = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
psd.L2_BUDGET_BYTES = int(os.environ.get("CT_L2", psd.L2_BUDGET_BYTES))   # intermediate budget (sets the batch)
g = torch.Generator(device="cuda"); g.manual_seed(3)
x = torch.randn(n, generator=g, device="cuda") * 24 + 5000
acc, nseg = psd.welch_sums(x, L); torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): acc, nseg = psd.welch_sums(x, L, shift=5000.0)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(f"n={n} L={L} nseg={nseg}: {ms:.3f} ms  {n/ms/1e6:.1f} Gsamples/s (new samples)  {4*n/ms/1e6:.0f} GB/s  frac {4*n/ms/1e6/6559.7:.3f}")
