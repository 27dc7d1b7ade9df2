#!/usr/bin/env python
"""Static SASS counts per kernel of bmfr_b200/libbmfr_b200.so (cuobjdump -sass) as a markdown table:
the instructions that show which hardware paths a kernel uses (TMA, mbarrier, packed fp32, redux, LDS vs generic LD)."""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "bmfr_b200/libbmfr_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
COLS = ["UTMALDG", "SYNCS", "FFMA2", "FADD2", "FMUL2", "CREDUX", "SHFL", "LDS", "LD.E", "LDG", "HMMA"]
rows, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        rows[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if cur and m:
        op = m.group(1)
        rows[cur]["instr"] += 1
        for c in COLS:
            if op == c or op.startswith(c + ".") or (c == "LD.E" and op.startswith("LD.E")):
                rows[cur][c] += 1
print(f"# SASS evidence (cuobjdump -sass {lib}, static counts; scripts/sass_evidence.py)\n")
print("| kernel | SASS instr | KB | " + " | ".join(c + {"UTMALDG": " (TMA)", "SYNCS": " (mbarrier)", "LD.E": " (generic)"}.get(c, "") for c in COLS) + " |")
print("|---|---|---|" + "---|" * len(COLS))
for k in sorted(rows):
    r = rows[k]
    print(f"| {k} | {r['instr']} | {r['instr'] * 16 / 1024:.1f} | " + " | ".join(str(r[c]) for c in COLS) + " |")
