"""Summarise an `ncu --page source --csv` export: executed instructions per opcode and the hottest stall sites.
usage: python tools/ncu_src_summary.py <source.csv> <particle_steps>"""
import collections
import csv
import sys

path, psteps = sys.argv[1], float(sys.argv[2])
rows = list(csv.reader(open(path)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
ops, stall_by_op = collections.Counter(), collections.Counter()
tot = 0
samples = collections.Counter()
reasons = collections.Counter()
rcols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
lines = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    src = r[ix["Source"]].strip()
    toks = src.split()
    op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
    op = op.split(".")[0]
    n = int(r[ix["Instructions Executed"]] or 0)
    s = int(r[ix["# Samples"]] or 0)
    ops[op] += n
    tot += n
    stall_by_op[op] += s
    for h in rcols:
        reasons[h] += int(r[ix[h]] or 0)
    lines.append((s, n, src))
print(f"warp instructions {tot:.3e}; lane-instr per particle-step {tot * 32 / psteps:.0f}")
print("-- executed by opcode (per particle-step, % of total) | stall samples %")
ssum = sum(stall_by_op.values())
for op, n in ops.most_common(28):
    print(f"{op:12s} {n * 32 / psteps:9.1f} {100 * n / tot:6.2f}%   | {100 * stall_by_op[op] / ssum:6.2f}%")
print("-- stall reasons (all samples)")
rs = sum(reasons.values())
for h, n in reasons.most_common(10):
    print(f"{h:28s} {100 * n / rs:6.2f}%")
print("-- hottest instructions by samples")
for s, n, src in sorted(lines, reverse=True)[:25]:
    print(f"{100 * s / ssum:6.2f}%  exec {n:12d}  {src[:100]}")
