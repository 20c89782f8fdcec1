#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: stall totals, instruction mix by opcode,
and the hottest SASS lines.  Usage: ncu_source_summary.py src.csv [top_n]"""
import csv
import sys
from collections import Counter

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path)))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) == len(hdr)]
stall_cols = [h for h in hdr if h.startswith("stall_")]
tot = Counter()
for r in body:
    for h in stall_cols:
        try:
            tot[h] += int(r[idx[h]])
        except ValueError:
            pass
allsamp = sum(tot.values())
print("stall samples:", allsamp)
for h, v in tot.most_common(12):
    print(f"  {h:24s} {v:8d} {100.0 * v / max(allsamp, 1):5.1f}%")
ops = Counter()
execd = Counter()
for r in body:
    src = r[idx["Source"]].strip()
    op = src.split()[0] if src else "?"
    if op.startswith("@"):
        op = src.split()[1]
    try:
        n = int(r[idx["Instructions Executed"]])
        s = int(r[idx["# Samples"]])
    except ValueError:
        continue
    ops[op.split(".")[0]] += s
    execd[op.split(".")[0]] += n
print("instructions executed by opcode (top):", sum(execd.values()))
for op, n in execd.most_common(22):
    print(f"  {op:12s} exec {n:12d} {100.0 * n / sum(execd.values()):5.1f}%   samples {ops[op]:7d}")
print("hottest lines:")
lines = sorted(body, key=lambda r: -int(r[idx["# Samples"]] or 0))[:top]
for r in lines:
    stalls = {h: int(r[idx[h]] or 0) for h in stall_cols}
    top2 = sorted(stalls.items(), key=lambda kv: -kv[1])[:2]
    print(f"  {r[idx['# Samples']]:>7s} {r[idx['Instructions Executed']]:>10s}  "
          f"{r[idx['Source']][:70]:70s} {top2}")
