#!/usr/bin/env python
"""Summarise `ncu --page source --csv` output: instruction mix and stall reasons of one kernel.
usage: ncu -i rep.ncu-rep --page source --csv --kernel-name K --launch-count 1 > k.csv; ncu_sass_summary.py k.csv"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
idx = {name: i for i, name in enumerate(hdr)}
stall_cols = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
ops, samples, stalls = collections.Counter(), collections.Counter(), collections.Counter()
tot = tot_s = 0
for r in rows[h + 1:]:
    if len(r) < len(hdr) or not r[idx["Instructions Executed"]].isdigit():
        continue
    sass = r[idx["Source"]].strip()
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", sass)
    op = m.group(2).split(".")[0] if m else sass[:10]
    n, s = int(r[idx["Instructions Executed"]]), int(r[idx["# Samples"]])
    ops[op] += n; samples[op] += s; tot += n; tot_s += s
    for c in stall_cols:
        stalls[c] += int(r[idx[c]])
print(f"total warp-instructions {tot}, samples {tot_s}, static SASS lines {len(rows) - h - 1}")
for op, n in ops.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 24):
    print(f"  {op:10s} {n:11d} {100 * n / tot:5.1f}%   stall samples {100 * samples[op] / max(tot_s, 1):5.1f}%")
print("stall reasons (% of samples):", {k: round(100 * v / max(tot_s, 1), 1) for k, v in stalls.most_common(9)})
