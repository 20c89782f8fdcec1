#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout -s KILL 420 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -5
echo "=== bench fused"
timeout -s KILL 200 python bench.py --no-cpu --no-e2e --no-online --no-deskew --steps 20 > gpurun_out/bench_f2_fused.json 2> gpurun_out/bench_f2_fused.err; python - <<'PY'
import json
for l in open("gpurun_out/bench_f2_fused.json"):
    if l.startswith("{"):
        d=json.loads(l); print("%.1f Gpts/s step %.3f ms k_decode %.3f ms" % (d["value"]/1e9, d["ms_per_step"], d["roofline"]["kernel_ms"]))
PY
bash scratch/run_prof.sh
