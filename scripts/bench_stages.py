"""Stage-by-stage timing of pipeline.TraceAnalyzer on a C2-style trace (device resident), for tuning."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cusumtools_b200 import detect, filters, pipeline, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 29
S = synth.CHIMERA_SETTINGS
raw = synth.device_trace(n, "cuda", seed=1234)
an = pipeline.TraceAnalyzer(n, S, 1e5, 8, threshold=5.0, hysteresis=1.0, baseline_block=1 << 20, baseline_min=4700.0,
                            baseline_max=5300.0, cusum_delta=400.0, cusum_h=10.0, fuse_stats=bool(int(os.environ.get('FUSE', '1'))), fused_count=bool(int(os.environ.get('FUSEC', '1'))))
for rep in range(4):
    marks = {}
    def hook(name):
        e = torch.cuda.Event(enable_timing=True); e.record(); marks[name] = (e, time.perf_counter())
    torch.cuda.synchronize(); hook("start")
    r = an.run(raw, stage_hook=hook)
    torch.cuda.synchronize()
    names = ["start", "median", "filter", "baseline", "detect", "cusum"]
    print("run", rep, "  ".join(f"{b} {marks[a][0].elapsed_time(marks[b][0]):.2f}ms(host {1e3*(marks[b][1]-marks[a][1]):.2f})" for a, b in zip(names, names[1:])),
          "events", len(r.events))
# the filter alone, back to back
med = r.median_codes
bl = detect.new_baseline(n, 1 << 20, 4700.0, 5300.0, raw.device)
for stats in (None, detect.stats_args(bl)):
    for _ in range(2): filters.dequant_filtfilt(raw, S, 1e5, 8, median_codes=med, out=an.y, workspace=an.filter_ws, stats=stats)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): filters.dequant_filtfilt(raw, S, 1e5, 8, median_codes=med, out=an.y, workspace=an.filter_ws, stats=stats)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"filter alone stats={'fused' if stats is not None else 'off'}: {ms:.3f} ms  {n/ms/1e6:.1f} Gs/s  frac {6*n/ms/1e6/6559.7:.3f}")
