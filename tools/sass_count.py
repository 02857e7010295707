#!/usr/bin/env python
"""Static SASS opcode counts of one kernel of libgradus_b200.so (default: gb200_trace_kernel<Kerr, ThinDisc>).

The step attempt of the trace kernel is straight-line code, so the static FP64 counts between the loop head and the
commit track the executed counts ncu reports; this lets the instruction budget of a variant be read on a box without a
GPU.  Also counts the three-register DFMAs (all operands distinct registers) with and without an operand-reuse flag.
usage: tools/sass_count.py [lib.so] [mangled-name-substring]"""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "gradus.jl_b200/csrc/libgradus_b200.so"
want = sys.argv[2] if len(sys.argv) > 2 else "gb200_trace_kernelILi0ELi1E"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, body = None, collections.defaultdict(list)
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    if cur and re.search(r"/\*[0-9a-f]{4,6}\*/", line):
        body[cur].append(line)
for name, lines in body.items():
    if want not in name:
        continue
    ops = collections.Counter()
    three, three_reuse = 0, 0
    for ln in lines:
        m = re.search(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)\s*(.*?);", ln)
        if not m:
            continue
        op = m.group(1).split(".")[0]
        ops[op] += 1
        if op == "DFMA":
            args = [a.strip() for a in m.group(2).split(",")][1:]
            regs = [re.sub(r"[-|~]|\.reuse", "", a) for a in args if re.match(r"[-|~]*R\d+", a)]
            if len(regs) == 3 and len(set(regs)) == 3:
                three += 1
                three_reuse += any(".reuse" in a for a in args)
    tot = sum(ops.values())
    fp64 = ops["DFMA"] + ops["DMUL"] + ops["DADD"] + ops["DSETP"]
    print(f"{name}: {tot} instructions, FP64 {fp64} (DFMA {ops['DFMA']} DMUL {ops['DMUL']} DADD {ops['DADD']} DSETP {ops['DSETP']}), "
          f"three-register DFMA {three} (with .reuse {three_reuse}), MUFU {ops['MUFU']}, LDL {ops['LDL']} STL {ops['STL']}, LDS {ops['LDS']} STS {ops['STS']}")
    print("  top:", ", ".join(f"{k} {v}" for k, v in ops.most_common(14)))
