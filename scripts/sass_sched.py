#!/usr/bin/env python
"""Static view of a SASS region (cuobjdump -sass text): opcode histogram, per-instruction stall counts /
wait masks decoded from the control bits (stall = bits 105..108, yield = 109, wbar = 110..112, rbar = 113..115,
wait mask = 116..121 of the 128-bit word, per /opt/skills/guides/B300_MICROARCH.md), and the issue time of
ONE warp running the region alone (fixed-latency stalls only).

    cuobjdump -sass -fun <mangled> file.o > k.sass
    python scripts/sass_sched.py k.sass 0x2650 0x4df0 [--list]
"""
import re
import sys
from collections import Counter

def parse(path):
    ins = []
    lines = open(path).read().splitlines()
    i = 0
    pat = re.compile(r"/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\* (0x[0-9a-f]{16}) \*/")
    hi = re.compile(r"^\s*/\* (0x[0-9a-f]{16}) \*/")
    while i < len(lines):
        m = pat.search(lines[i])
        if m and i + 1 < len(lines):
            h = hi.match(lines[i + 1])
            if h:
                addr = int(m.group(1), 16)
                text = m.group(2).strip()
                w1 = int(h.group(1), 16)
                ctrl = w1 >> 41
                stall = ctrl & 0xf
                yld = (ctrl >> 4) & 1
                wbar = (ctrl >> 5) & 7
                rbar = (ctrl >> 8) & 7
                wait = (ctrl >> 11) & 0x3f
                ins.append((addr, text, stall, yld, wbar, rbar, wait))
                i += 2
                continue
        i += 1
    return ins

def opcode(text):
    t = text.split()
    k = 0
    while k < len(t) and t[k].startswith("@"):
        k += 1
    return t[k].split(".")[0] if k < len(t) else "?"

def main():
    path = sys.argv[1]
    lo = int(sys.argv[2], 16) if len(sys.argv) > 2 else 0
    hi = int(sys.argv[3], 16) if len(sys.argv) > 3 else 1 << 62
    ins = [x for x in parse(path) if lo <= x[0] <= hi]
    hist = Counter(opcode(x[1]) for x in ins)
    tot = len(ins)
    cyc = sum(max(1, x[2]) for x in ins)
    print(f"{tot} instructions, sum of stall fields {cyc} cycles (one warp alone, no scoreboard waits)")
    for op, c in hist.most_common():
        print(f"  {op:12s} {c:5d}")
    if "--list" in sys.argv:
        for a, t, s, y, wb, rb, wm in ins:
            print(f"{a:06x} s{s:2d} {'Y' if not y else ' '} w{wb} r{rb} m{wm:02x}  {t}")

if __name__ == "__main__":
    main()


def simulate(ins, nwarps, iters=6):
    """Issue-level model of `nwarps` warps looping over the region on ONE scheduler: one instruction per cycle,
    a warp waits its own stall field after each issue, and the FMA / ALU pipes accept one warp instruction every
    2 cycles (B300_MICROARCH: rt_SMSP = 2).  Memory latencies are ignored: an upper bound on throughput."""
    FMA = {"FFMA2", "FFMA", "FMUL", "FMUL2", "FADD", "FADD2", "IMAD", "HFMA2", "DFMA"}
    ALU = {"IADD3", "LOP3", "SHF", "PRMT", "FMNMX", "FMNMX3", "ISETP", "FSETP", "SEL", "FSEL", "LEA", "MOV", "VIADD", "PLOP3",
           "IABS", "VIMNMX", "F2I", "I2F", "POPC", "FLO", "BREV"}
    n = len(ins)
    pc = [0] * nwarps
    ready = [w for w in range(nwarps)]          # staggered start
    done = [0] * nwarps
    pipe_free = {"fma": 0, "alu": 0}
    t = 0
    last = 0
    total = n * iters
    while min(done) < total:
        issued = False
        for k in range(nwarps):
            w = (last + 1 + k) % nwarps
            if done[w] >= total or ready[w] > t:
                continue
            op = opcode(ins[pc[w]][1])
            pipe = "fma" if op in FMA else ("alu" if op in ALU else None)
            if pipe and pipe_free[pipe] > t:
                continue
            if pipe:
                pipe_free[pipe] = t + 2
            ready[w] = t + max(1, ins[pc[w]][2])
            pc[w] = (pc[w] + 1) % n
            done[w] += 1
            last = w
            issued = True
            break
        t += 1
    return t / (iters * nwarps)


if __name__ == "__main__" and "--sim" in sys.argv:
    path = sys.argv[1]
    lo, hi = int(sys.argv[2], 16), int(sys.argv[3], 16)
    ins = [x for x in parse(path) if lo <= x[0] <= hi]
    for nw in (1, 2, 3, 4):
        print(f"  {nw} warps/scheduler: {simulate(ins, nw):.1f} cycles per loop iteration per warp-iteration")
