#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout -s KILL 300 python -m pytest tests/test_gpu_deskew.py tests/test_gpu_parity.py -x -q 2>&1 | tail -3
timeout -s KILL 200 python bench.py --no-cpu --no-e2e --no-online --steps 10 > gpurun_out/bench_r1z_dsk.json 2>/dev/null
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_r1z_dsk.json") if l.startswith("{")][-1])
print("value %.1f G; deskew %.1f G ms %.3f kdec %.3f" % (d["value"]/1e9, d["deskew_per_point"]["points_per_s_per_gpu"]/1e9, d["deskew_per_point"]["ms_per_step"], d["deskew_per_point"]["k_decode_ms"]))
PY
