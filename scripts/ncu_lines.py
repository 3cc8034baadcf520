#!/usr/bin/env python
"""Per-CUDA-source-line executed warp-instructions from `ncu --page source --csv --print-source cuda,sass`.
usage: scripts/ncu_lines.py src_cuda.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = []; tot = 0; fname = ''
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path': fname = r[1].split('/')[-1]; continue
    if len(r) >= 8 and r[0].isdigit():
        try: n = int(r[7])
        except ValueError: continue
        out.append((n, fname, int(r[0]), r[1].strip()[:100], int(r[4]) if r[4].isdigit() else 0)); tot += n
stall = sum(o[4] for o in out)
print("total warp-instructions", tot)
for n, f, ln, src, st in sorted(out, reverse=True)[:top]:
    print(f"{100*n/tot:5.1f}%  stall {100*st/max(stall,1):5.1f}%  {f}:{ln:<4d} {src}")
