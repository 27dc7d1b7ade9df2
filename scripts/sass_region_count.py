#!/usr/bin/env python
"""Static SASS count per source region of one kernel, following inline chains.

usage: sass_region_count.py cubin kernel_substring file:lo-hi[:name] ...
Disassembles with `nvdisasm -gi` (needs -lineinfo), attributes every instruction to the FIRST region (in argument order)
that contains any link of its inline chain, prints counts and the opcode mix per region (DUMP=name: also its instructions).  For straight-line code the
static count of a region is its dynamic count per warp — the number to look at before spending GPU time."""
import collections
import re
import subprocess
import sys

cubin, kernel = sys.argv[1], sys.argv[2]
regions = []
for a in sys.argv[3:]:
    parts = a.split(":")
    lo, hi = map(int, parts[1].split("-"))
    regions.append((parts[0], lo, hi, parts[2] if len(parts) > 2 else a))
txt = subprocess.run(["nvdisasm", "-gi", cubin], capture_output=True, text=True).stdout.splitlines()
import os
dump = os.environ.get("DUMP")
inside, chain, pending = False, [], []
count, ops = collections.Counter(), collections.defaultdict(collections.Counter)
for l in txt:
    if l.startswith("\t.section\t.text."):
        inside = kernel in l
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
    if m:
        pending.append((m.group(1).split("/")[-1], int(m.group(2))))
        if m.group(3):
            pending.append((m.group(3).split("/")[-1], int(m.group(4))))
        continue
    mi = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\w+\s+)?([A-Z0-9_.]+)", l)
    if mi:
        if pending:
            chain, pending = pending, []
        reg = "other"
        for f, lo, hi, name in regions:
            if any(cf == f and lo <= cl <= hi for cf, cl in chain):
                reg = name
                break
        count[reg] += 1
        if dump == reg:
            print(f"{chain[0][0]}:{chain[0][1]:<5d}", l.split("*/", 1)[1].strip()[:90])
        ops[reg][mi.group(2).split(".")[0]] += 1
for reg, n in count.most_common():
    print(f"{n:6d} {reg}: " + " ".join(f"{o}:{c}" for o, c in ops[reg].most_common(14)))
