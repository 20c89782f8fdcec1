#!/bin/bash
# round 2: full-size parity tests, default bench (with the configs[3] leg), 1-h headline at N=1
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu > gpurun_out/r2d_fullsize.log 2>&1
echo "fullsize exit $?" >> gpurun_out/r2d_fullsize.log
tail -15 gpurun_out/r2d_fullsize.log
timeout 1200 python bench.py > gpurun_out/bench_r2d.json 2> gpurun_out/bench_r2d.err
echo "bench exit $?"; tail -c 800 gpurun_out/bench_r2d.err
timeout 900 python bench.py --recording-hours 1 --steps 5 > gpurun_out/bench_r2d_rec1h.json 2> gpurun_out/bench_r2d_rec1h.err
echo "rec exit $?"; tail -c 800 gpurun_out/bench_r2d_rec1h.err
