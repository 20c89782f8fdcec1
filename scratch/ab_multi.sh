#!/bin/bash
# A/B/C... of several builds of the library on the SAME box: scratch/ab_multi.sh rounds lib1.so lib2.so ...
cd "$GRAFT_REPO_ROOT"
R=$1; shift
cp veloslam_b200/libveloslam_b200.so /tmp/orig.so
for r in $(seq 1 $R); do
  for f in "$@"; do
    cp $f veloslam_b200/libveloslam_b200.so
    timeout 200 python bench.py --no-cpu --no-e2e --no-online --no-deskew --no-single-pass --no-parity --no-facade --no-hdl32 --recording-leg-hours 0 --online-udp-seconds 0 --steps 20 > /tmp/ab.json 2>/tmp/ab.err
    python - <<PY
import json
ok=False
for l in open("/tmp/ab.json"):
    if l.startswith("{"):
        d=json.loads(l); ok=True
        print("AB $f round $r: %.1f Gpts/s step %.3f ms k_decode %.3f ms frac_whole %.3f" % (d["value"]/1e9, d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac_whole_step"]))
if not ok: print("AB $f round $r FAILED", open("/tmp/ab.err").read()[-300:])
PY
  done
done
cp /tmp/orig.so veloslam_b200/libveloslam_b200.so
