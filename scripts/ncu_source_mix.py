#!/usr/bin/env python
"""Dynamic SASS instruction mix + top stall sites from an ncu source page.
usage: ncu -i rep --page source --csv > src.csv ; scripts/ncu_source_mix.py src.csv [samples]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
nsamp = float(sys.argv[2]) if len(sys.argv) > 2 else None
hdr = rows[1]
iS, iE, iSt = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('Warp Stall Sampling (All Samples)')
mix = collections.Counter(); stall = collections.Counter(); tot = 0; tots = 0
lines = []
for r in rows[2:]:
    if len(r) <= iE:
        if r and r[0] == 'Kernel Name': break      # only the first launch of the page
        continue
    if r[iE] == 'Instructions Executed': continue
    src = r[iS].strip(); n = int(r[iE] or 0); s = int(r[iSt] or 0)
    parts = src.split()
    op = parts[1] if parts and parts[0].startswith('@') and len(parts) > 1 else (parts[0] if parts else '?')
    op = op.split('.')[0]
    mix[op] += n; stall[op] += s; tot += n; tots += s
    lines.append((s, n, src))
print(f"total warp-instructions {tot}" + (f"  = {tot*32/nsamp:.1f} thread-instr/sample" if nsamp else ""))
for op, n in mix.most_common(22):
    print(f"  {op:12s} {n:12d} {100*n/tot:5.1f}%   stall-samples {100*stall[op]/max(tots,1):5.1f}%" + (f"   {n*32/nsamp:6.2f}/sample" if nsamp else ""))
print("top stall sites:")
for s, n, src in sorted(lines, reverse=True)[:25]:
    print(f"  {100*s/max(tots,1):5.2f}%  exec {n:9d}  {src[:90]}")
