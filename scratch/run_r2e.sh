#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2e_all_tests.log 2>&1
echo "all tests exit $?" >> gpurun_out/r2e_all_tests.log
tail -4 gpurun_out/r2e_all_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2e_smoke.log 2>&1; tail -2 gpurun_out/r2e_smoke.log
timeout 1500 python bench.py > gpurun_out/bench_r2e.json 2> gpurun_out/bench_r2e.err
echo "bench exit $?"; tail -c 600 gpurun_out/bench_r2e.err
