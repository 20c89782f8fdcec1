#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
bash scratch/ab_multi.sh 2 scratch/ab/c0.so scratch/ab/c1.so
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2x_all_tests.log 2>&1
echo "all tests exit $?"; tail -3 gpurun_out/r2x_all_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
