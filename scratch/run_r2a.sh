#!/bin/bash
# round 2, first GPU pass: new layout/facade tests, full GPU suite, bench
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt 2>&1
nproc >> gpurun_out/r2a_smi.txt
timeout 900 python -m pytest tests/test_gpu_layout.py tests/test_gpu_facade.py -x -q -m gpu > gpurun_out/r2a_new_tests.log 2>&1
echo "new tests exit $?" >> gpurun_out/r2a_new_tests.log
tail -5 gpurun_out/r2a_new_tests.log
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2a_all_tests.log 2>&1
echo "all tests exit $?" >> gpurun_out/r2a_all_tests.log
tail -5 gpurun_out/r2a_all_tests.log
timeout 900 python bench.py > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err
echo "bench exit $?"
tail -c 600 gpurun_out/bench_r2a.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2a_ref.json 2> gpurun_out/bench_r2a_ref.err
echo "ref exit $?"
