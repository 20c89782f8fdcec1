#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
LEAN="--no-e2e --no-cpu --no-online --no-deskew --no-single-pass --no-parity --no-facade --no-hdl32 --recording-leg-hours 0 --online-udp-seconds 0"
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/r2y_launches.csv python bench.py --steps 2 --warmup 1 $LEAN > gpurun_out/r2y_launches_bench.log 2>&1
echo "launch list exit $?"
timeout 900 python bench.py > gpurun_out/bench_r2y.json 2> gpurun_out/bench_r2y.err
echo "bench exit $?"; tail -c 300 gpurun_out/bench_r2y.err
