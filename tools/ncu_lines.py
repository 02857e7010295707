#!/usr/bin/env python
"""Join an .ncu-rep's per-instruction counters (source page, SASS) with the line table of the cubin (nvdisasm -g) and report
executed warp instructions / stall samples per source region of gb200_trace_kernel: step attempt (RHS, stage combinations,
error estimate + controller, event test, commit), event scan, service pass.  usage: tools/ncu_lines.py rep cubin [kernel-substring]"""
import collections, csv, io, re, subprocess, sys
rep, cubin = sys.argv[1], sys.argv[2]
want = sys.argv[3] if len(sys.argv) > 3 else "gb200_trace_kernelILi0ELi1E"
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if ".text." in l and want in l and "section" in l)
lines, cur = [], None
for l in dis[start + 1:]:
    if l.startswith("//--------------------- .text."):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]+\*/\s+\S", l):
        lines.append(cur)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = {h: i for i, h in enumerate(rows[1])}
ins = [r for r in rows[2:] if len(r) > hdr["Instructions Executed"]]
assert abs(len(ins) - len(lines)) < 8, (len(ins), len(lines))
per = collections.defaultdict(lambda: [0, 0, 0])
for r, ln in zip(ins, lines):
    op = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[hdr["Source"]]).group(2).split(".")[0]
    e, s = int(r[hdr["Instructions Executed"]]), int(r[hdr["# Samples"]])
    per[ln][0] += e; per[ln][1] += s
    if op in ("DFMA", "DMUL", "DADD", "DSETP"):
        per[ln][2] += e
tot = sum(v[0] for v in per.values()); tots = sum(v[1] for v in per.values()); totf = sum(v[2] for v in per.values())
print(f"executed warp instructions {tot}, FP64 {totf} ({100*totf/tot:.1f} %), samples {tots}")
byfile = collections.defaultdict(lambda: [0, 0, 0])
for (f, l), v in per.items():
    for k in range(3): byfile[f][k] += v[k]
for f, v in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
    print(f"  {f:32s} {100*v[0]/tot:6.2f} % instr  {100*v[2]/max(totf,1):6.2f} % of FP64  {100*v[1]/tots:6.2f} % samples")
print("top source lines:")
for (f, l), v in sorted(per.items(), key=lambda kv: -kv[1][0])[:45]:
    print(f"  {f}:{l:<5d} {100*v[0]/tot:6.2f} % instr  {100*v[2]/max(totf,1):6.2f} % of FP64  {100*v[1]/tots:6.2f} % samples")
