#!/bin/bash
# first GPU run of the single-pass pipeline: parity tests, then A/B bench against the two-pass one
cd $GRAFT_REPO_ROOT
timeout -s KILL 420 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -25
echo "=== rest"
timeout -s KILL 600 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_parity.py 2>&1 | tail -15
echo "=== bench fused"
timeout -s KILL 200 python bench.py --no-cpu --no-e2e --no-online --no-deskew --steps 20 > gpurun_out/bench_f1_fused.json 2> gpurun_out/bench_f1_fused.err; tail -c 600 gpurun_out/bench_f1_fused.json
echo "=== bench two-pass"
VELOSLAM_TWO_PASS=1 timeout -s KILL 200 python bench.py --no-cpu --no-e2e --no-online --no-deskew --steps 20 > gpurun_out/bench_f1_two.json 2> gpurun_out/bench_f1_two.err; tail -c 600 gpurun_out/bench_f1_two.json
