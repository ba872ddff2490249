#!/usr/bin/env python3
"""Summarise an .ncu-rep (read on the CPU box): key raw metrics + instruction / stall shares by
source line. Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 45
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum", "sm__inst_executed_pipe_lsu.sum.pct",
        "sm__inst_executed_pipe_alu.sum.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "launch__shared_mem_per_block_dynamic", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "lts__t_sectors_op_write.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__grid_size", "launch__block_size"]
for i, h in enumerate(hdr):
    if any(h == w or h.startswith(w) for w in want) and "per_second" not in h and "pct_of_peak_sustained_elapsed" not in h.replace("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "").replace("sm__throughput.avg.pct_of_peak_sustained_elapsed", ""):
        print(f"{h:75s} {rows[1][i]:12s} {[r[i] for r in rows[2:]]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
agg, cur, hdr = {}, None, None
for r in csv.reader(io.StringIO(src)):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if r[0] in ("", "Function Name"):
        continue
    try:
        line, inst, samp = int(r[0]), int(r[hdr.index("Instructions Executed")]), int(r[hdr.index("# Samples")])
    except Exception:
        continue
    a = agg.setdefault((cur, line), [0, 0, r[1][:100]])
    a[0] += inst
    a[1] += samp
tot = sum(a[0] for a in agg.values()) or 1
tots = sum(a[1] for a in agg.values()) or 1
byfile = {}
for (f, _), a in agg.items():
    b = byfile.setdefault(f, [0, 0])
    b[0] += a[0]
    b[1] += a[1]
print("instructions (warp-level):", tot, " samples:", tots)
print({f: (round(100 * b[0] / tot, 1), round(100 * b[1] / tots, 1)) for f, b in byfile.items() if b[0] * 200 > tot})
for (f, l), a in sorted(agg.items(), key=lambda kv: -(kv[1][0] / tot + kv[1][1] / tots))[:top]:
    print(f"{f:16s}:{l:4d} inst={100 * a[0] / tot:5.2f}% samp={100 * a[1] / tots:5.2f}%  {a[2]}")
