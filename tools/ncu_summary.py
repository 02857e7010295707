#!/usr/bin/env python
"""Summarise an .ncu-rep (raw + source pages) for the trace kernel: key metrics, stall mix, opcode mix.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [out.txt]"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
keep = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.per_cycle_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed"]
print(f"# {rep}", file=out)
for h, u, v in zip(hdr, units, vals):
    if h in keep or ("smsp__average_warps_issue_stalled" in h and float(v or 0) > 0.01):
        print(f"{h} [{u}] = {v}", file=out)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
tot_s = tot_i = 0
op_s, op_i = collections.Counter(), collections.Counter()
for r in rows[2:]:
    try:
        s = int(r[ix["# Samples"]]); ie = int(r[ix["Instructions Executed"]])
    except Exception:
        continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ix["Source"]])
    op = m.group(2).split(".")[0] if m else "?"
    tot_s += s; tot_i += ie; op_s[op] += s; op_i[op] += ie
print(f"static SASS instructions: {len(rows) - 2} ({(len(rows) - 2) * 16 / 1024:.1f} KB); executed warp instructions: {tot_i}", file=out)
print("opcode mix (share of executed warp instructions / share of stall samples):", file=out)
for op, v in op_i.most_common(16):
    print(f"  {op:8s} {100 * v / tot_i:6.2f}% instr {100 * op_s[op] / max(tot_s, 1):6.2f}% samples", file=out)
