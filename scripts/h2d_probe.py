"""Pinned host -> device copy bandwidth on this box (sizes the e2e pipeline)."""
import time, torch
for mb in (64, 512, 4768):
    n = mb << 20
    h = torch.empty(n, dtype=torch.uint8, pin_memory=True); h.fill_(1)
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    d.copy_(h, non_blocking=True); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    e0.record()
    for _ in range(3): h.copy_(d, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1) / 3
    print(f"{mb} MiB: H2D {n/ms/1e6:.1f} GB/s  D2H {n/ms2/1e6:.1f} GB/s")
