#!/usr/bin/env python
"""Pretty-print the interesting fields of a bench.py JSON line."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(f"value {d['value']:.1f} {d['unit']}  ms/frame {d.get('ms_per_frame', 0):.4f}  launches {d.get('gpu_launches')}  clocks {d.get('clocks')}")
for k, v in (d.get("kernels") or {}).items():
    if "achieved_gbs" in v:
        print(f"  {k:28s} {v['ms'] * 1e3:8.1f} us  {v['achieved_gbs']:7.0f} GB/s  {100 * v['frac']:5.1f}% of peak" +
              (f"   busy {v['busy_ms'] * 1e3:6.1f} us {100 * v['frac_busy']:5.1f}%" if "busy_ms" in v else ""))
    else:
        print(f"  {k:28s} {v['ms'] * 1e3:8.1f} us")
if d.get("e2e"):
    print("  e2e", round(d["e2e"]["value"], 1), d["e2e"]["unit"])
if d.get("cpu_baseline"):
    print("  cpu", round(d["cpu_baseline"]["value"], 2), d["cpu_baseline"]["kind"], d["cpu_baseline"]["cores"], "cores")
for k in ("in_order", "sustained"):
    if d.get(k):
        print(f"  {k}", round(d[k]["value"], 1), d[k].get("unit"), {a: b for a, b in d[k].items() if a in ("steps", "seconds", "ms_per_frame")})
if d.get("parity"):
    print("  parity", d["parity"])
