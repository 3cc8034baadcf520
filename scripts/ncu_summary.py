#!/usr/bin/env python
"""Summarise an ncu report (.ncu-rep) into the handful of numbers DESIGN.md / profiles/ cite.
usage: scripts/ncu_summary.py report.ncu-rep [regex-filter]"""
import csv, io, re, subprocess, sys

KEYS = [r"gpu__time_duration\.sum", r"dram__bytes_read\.sum$", r"dram__bytes_write\.sum$",
        r"gpu__dram_throughput\.avg\.pct", r"sm__throughput\.avg\.pct", r"sm__warps_active\.avg\.pct",
        r"launch__registers_per_thread", r"launch__occupancy_limit", r"launch__grid_size", r"launch__block_size",
        r"launch__shared_mem_per_block_dynamic", r"sm__inst_executed\.sum$", r"smsp__inst_executed\.avg\.per_cycle_active",
        r"smsp__issue_active\.avg\.pct", r"sm__cycles_elapsed\.avg$", r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum$",
        r"sm__inst_executed_pipe_(fma|alu|fp64|xu|lsu|fmaheavy|fmalite)[a-z_]*\.sum$", r"sm__pipe_(fma|alu|fp64)_cycles_active\.avg\.pct",
        r"smsp__average_warps_issue_stalled_.*_per_issue_active\.ratio", r"smsp__warp_issue_stalled_.*_per_warp_active\.pct",
        r"lts__t_bytes\.sum$", r"lts__t_sector_hit_rate\.pct", r"l1tex__t_sector_hit_rate\.pct", r"smsp__cycles_active\.avg$",
        r"sm__warps_active\.avg\.per_cycle_active", r"achieved_occupancy", r"smsp__thread_inst_executed_per_inst_executed"]


def main():
    rep = sys.argv[1]
    flt = sys.argv[2] if len(sys.argv) > 2 else None
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print(f"== {name[:100]}  grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}")
        for i, h in enumerate(hdr):
            short = h.split(".", 2)[-1] if h.count(".") >= 2 and h.split(".")[0].isupper() else h
            if any(re.search(k, h) for k in KEYS) and (flt is None or re.search(flt, h)):
                v = r[i]
                if v not in ("", "0", "n/a"):
                    print(f"   {h} [{units[i]}] = {v}")


if __name__ == "__main__":
    main()
