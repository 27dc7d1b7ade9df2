#!/usr/bin/env python
"""Per-CUDA-source-line instruction / stall-sample shares from
`ncu -i rep --page source --print-source cuda,sass --csv --kernel-name K --launch-count 1`."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file, hdr, idx, out = None, None, None, []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = r
        idx = {n: i for i, n in enumerate(hdr)}
    elif hdr and len(r) >= len(hdr) and r[0].isdigit():
        try:
            out.append((int(r[idx["Instructions Executed"]]), int(r[idx["# Samples"]]), cur_file, int(r[0]), r[1].strip()[:100]))
        except ValueError:
            pass
tot, tots = sum(o[0] for o in out), sum(o[1] for o in out)
print(f"total warp-instructions {tot}  samples {tots}")
for n, s, f, ln, src in sorted(out, reverse=True)[:top]:
    print(f"{100 * n / tot:5.1f}% inst {100 * s / max(tots, 1):5.1f}% smp  {f}:{ln:<4d} {src}")
