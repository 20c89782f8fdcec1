#!/bin/bash
# round-end style run at HEAD: whole GPU suite, smoke, bench (both arms)
cd $GRAFT_REPO_ROOT
timeout -s KILL 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout -s KILL 200 python __graft_entry__.py smoke 2>&1 | tail -1
timeout -s KILL 400 python bench.py > gpurun_out/bench_r1z.json 2> gpurun_out/bench_r1z.err; tail -c 300 gpurun_out/bench_r1z.json
timeout -s KILL 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1z_ref.json 2> gpurun_out/bench_r1z_ref.err; cut -c1-200 gpurun_out/bench_r1z_ref.json
