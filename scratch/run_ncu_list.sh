#!/bin/bash
# ncu launch lists (per-launch durations, cold cache, serialised) for two builds of the library
cd $GRAFT_REPO_ROOT
cp veloslam_b200/libveloslam_b200.so /tmp/orig.so
for v in base new; do
  cp scratch/lib_$v.so veloslam_b200/libveloslam_b200.so
  timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1z_$v.csv python bench.py --no-cpu --no-e2e --no-online --no-deskew --steps 2 --warmup 3 > /dev/null 2>&1
  python - <<PY
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/launches_r1z_$v.csv")) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
d=collections.defaultdict(list)
for r in rows[1:]:
    d[r[ki][:40]].append(float(r[vi].replace(",","")))
for k,v in d.items(): print("$v", k, len(v), "median %.1f us" % (sorted(v)[len(v)//2]/1e3))
PY
done
cp /tmp/orig.so veloslam_b200/libveloslam_b200.so
