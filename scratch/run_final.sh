#!/bin/bash
# round-end style run at HEAD: whole GPU suite, smoke, bench (both arms), ncu launch list
cd $GRAFT_REPO_ROOT
timeout -s KILL 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout -s KILL 200 python __graft_entry__.py smoke 2>&1 | tail -2
timeout -s KILL 400 python bench.py > gpurun_out/bench_r1z.json 2> gpurun_out/bench_r1z.err; tail -c 400 gpurun_out/bench_r1z.json
timeout -s KILL 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1z_ref.json 2> gpurun_out/bench_r1z_ref.err; cut -c1-600 gpurun_out/bench_r1z_ref.json
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1z.csv python bench.py --no-cpu --no-e2e --no-online --no-deskew --steps 2 --warmup 1 > /dev/null 2>&1
grep -c k_decode gpurun_out/launches_r1z.csv
