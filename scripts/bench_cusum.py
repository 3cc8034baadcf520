"""Kernel-only timing of ct_cusum_batch on config C3 (pre-extracted events)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cusumtools_b200 import cusum, synth
ne = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
x, offs, nlev = synth.c3_events_device(ne, "cuda")
w0 = offs[:-1].contiguous(); w1 = offs[1:].contiguous()
t = cusum.cusum_levels(x, w0, w1, delta=400.0, h=10.0); torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters): t = cusum.cusum_levels(x, w0, w1, delta=400.0, h=10.0)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
n = x.numel()
ok = (t.n_levels.to(torch.int64) == nlev + 2).float().mean().item()
print(f"events={ne} samples={n}: {ms:.3f} ms  {n/ms/1e6:.1f} Gsamples/s  {ne/ms*1e3/1e6:.2f} Mevents/s  {4*n/ms/1e6:.0f} GB/s  frac {4*n/ms/1e6/6559.7:.3f}  levels-recovered {ok:.3f}")
