#!/bin/bash
# A/B of two builds of the library on the SAME box: scratch/ab.sh libA.so libB.so [rounds]
# (bench.py loads veloslam_b200/libveloslam_b200.so; the variants are copied over it in turn)
A=$1; B=$2; R=${3:-3}
cp veloslam_b200/libveloslam_b200.so /tmp/orig.so
for r in $(seq 1 $R); do
  for v in A B; do
    eval f=\$$v
    cp $f veloslam_b200/libveloslam_b200.so
    timeout 120 python bench.py --no-cpu --no-e2e --no-online --no-deskew --no-single-pass --steps 20 > /tmp/ab.json 2>/dev/null
    python - <<PY
import json
for l in open("/tmp/ab.json"):
    if l.startswith("{"):
        d=json.loads(l); print("AB $v round $r: %.1f Gpts/s step %.3f ms k_decode %.3f ms" % (d["value"]/1e9, d["ms_per_step"], d["roofline"]["kernel_ms"]))
PY
  done
done
cp /tmp/orig.so veloslam_b200/libveloslam_b200.so
