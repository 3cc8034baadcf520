"""Stage-by-stage timing on a C2-style trace (device resident), for tuning."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cusumtools_b200 import cusum, detect, filters, pipeline, synth, _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 29
S = synth.CHIMERA_SETTINGS
raw = synth.device_trace(n, "cuda", seed=1234)
def T(fn, reps=3):
    r = fn(); torch.cuda.synchronize(); r = None
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(reps):
        r = None
        r = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, (time.perf_counter() - t0) * 1e3 / reps, r
mask = filters.chimera_bitmask(S)
ms, wall, med = T(lambda: filters.code_median(raw, mask)); print(f"median   {ms:8.3f} ms (wall {wall:.3f})")
ms, wall, y = T(lambda: filters.dequant_filtfilt(raw, S, 1e5, 8, median_codes=med)); print(f"filter   {ms:8.3f} ms (wall {wall:.3f})  {n/ms/1e6:.1f} Gs/s")
ms, wall, bl = T(lambda: detect.baseline_blocks(y, 1 << 20, 4700.0, 5300.0)); print(f"baseline {ms:8.3f} ms (wall {wall:.3f})")
bl = bl.with_thresholds(5.0, 1.0)
ms, wall, ev = T(lambda: detect.detect_events(y, bl)); print(f"detect   {ms:8.3f} ms (wall {wall:.3f})  events {len(ev)}")
ms, wall, w = T(lambda: detect.event_windows(ev.starts, ev.ends, n, 100, 8, 100000)); print(f"windows  {ms:8.3f} ms (wall {wall:.3f})")
w0, w1, typ = w
ns = int((w1 - w0).sum().item())
ms, wall, lv = T(lambda: cusum.cusum_levels(y, w0, w1, delta=400.0, h=10.0, types=typ)); print(f"cusum    {ms:8.3f} ms (wall {wall:.3f})  {ns/ms/1e6:.1f} Gs/s over {ns} event samples")
print("types", torch.bincount(typ).tolist(), "levels", torch.bincount(lv.n_levels).tolist(), "overflow", int(lv.overflow.sum()))
