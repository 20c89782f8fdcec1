#!/usr/bin/env python
"""Attribute the stall samples of an `ncu --page source --csv` dump (SASS rows) to CUDA source
lines, using `nvdisasm --print-line-info` of the cubin the kernel came from.
Usage: ncu_lines.py sass.csv disasm.txt 'kernel symbol substring' [top_n]
  disasm.txt: cuobjdump -xelf all lib.so; nvdisasm --print-line-info x.cubin > disasm.txt"""
import collections
import csv
import os
import re
import sys

sass_csv, dis_path, sym = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 50
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
cur, seq, inside = None, {}, False
for ln in open(dis_path):
    if ln.startswith("//-----"):
        if inside:
            break
        inside = (".text." in ln) and (sym in ln)
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        seq[int(m.group(1), 16)] = (cur, m.group(2).strip())
rows = list(csv.reader(open(sass_csv)))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) == len(hdr)]
num = lambda a: int(a, 16) if a.startswith("0x") else int(a)
base = num(body[0][idx["Address"]])
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
per, ins, why = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
for r in body:
    off = num(r[idx["Address"]]) - base
    if off not in seq:
        continue
    c, _ = seq[off]
    per[c] += int(r[idx["# Samples"]] or 0)
    ins[c] += int(r[idx["Instructions Executed"]] or 0)
    for h in stall_cols:
        why[c][h] += int(r[idx[h]] or 0)
tot = sum(per.values())
print("samples", tot, "warp instructions", sum(ins.values()))
text = {}
for f in ("vs_kernels.cuh", "vs_device.cuh"):
    text[f] = open(os.path.join(ROOT, "veloslam_b200", "csrc", f)).read().split("\n")
for (f, l), s in per.most_common(top):
    t = text[f][l - 1].strip()[:80] if f in text else ""
    w = ",".join("%s %d" % (k[6:], v) for k, v in why[(f, l)].most_common(2))
    print("%6d %5.1f%% inst %9d  %s:%d  %-80s [%s]" % (s, 100.0 * s / tot, ins[(f, l)], f, l, t, w))
