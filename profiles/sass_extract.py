#!/usr/bin/env python
"""SASS evidence for the sm_100a kernels of libveloslam_b200.so: per kernel the count of the
mnemonics that matter -- TMA bulk copies (UBLKCP.S.G global->shared, UBLKCP.G.S shared->global),
mbarrier operations (SYNCS.*), FP64 arithmetic (DMUL / DADD separate: no FMA contraction on the
reference-parity path; DFMA only in the per-point deskew extension and the geodesy kernel),
shared / global accesses, shuffles, votes, reductions.
Usage: python profiles/sass_extract.py [lib.so] > profiles/r2_sass_extract.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "veloslam_b200/libveloslam_b200.so"
sass = subprocess.check_output(["cuobjdump", "-sass", lib], text=True)
names = {}
counts = collections.OrderedDict()
fn = None
keep = ("DMUL", "DADD", "DFMA", "DSETP", "LDS", "STS", "LDG", "STG", "LDC", "SHFL", "BAR", "REDUX",
        "ATOMG", "ATOMS", "RED", "VOTE", "ELECT", "MUFU", "POPC", "CCTL", "MEMBAR", "FENCE")
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        counts[fn] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m or fn is None:
        continue
    op = m.group(1)
    base = op.split(".")[0]
    counts[fn]["(all)"] += 1
    if base == "UBLKCP":
        counts[fn][".".join(op.split(".")[:3])] += 1
    elif base == "SYNCS":
        counts[fn][".".join(op.split(".")[:2])] += 1
    elif base in keep:
        counts[fn][base] += 1
dem = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.split("\n")
for (f, c), d in zip(counts.items(), dem):
    print(f"{d}   [{c['(all)']} SASS instructions]")
    for k in sorted(k for k in c if k != "(all)"):
        print(f"    {k:28s} {c[k]}")
print()
print("--- first UBLKCP / SYNCS lines of k_decode<0,0,0> ---")
on = False
n = 0
for line in sass.splitlines():
    if "Function :" in line:
        on = "k_decodeILi0ELi0ELi0E" in line
    elif on and re.search(r"UBLKCP|SYNCS", line) and n < 24:
        print("   ", re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", line).strip())
        n += 1
