#!/bin/bash
# final-HEAD validation: whole GPU suite, smoke, launch list of the bench command, full bench
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2u_all_tests.log 2>&1
echo "all tests exit $?" >> gpurun_out/r2u_all_tests.log
tail -4 gpurun_out/r2u_all_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2u_smoke.log 2>&1; tail -2 gpurun_out/r2u_smoke.log
LEAN="--no-e2e --no-cpu --no-online --no-deskew --no-single-pass --no-parity --no-facade --no-hdl32 --recording-leg-hours 0 --online-udp-seconds 0"
timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/r2u_launches.csv python bench.py --steps 2 --warmup 1 $LEAN > gpurun_out/r2u_launches_bench.log 2>&1
echo "launch list exit $?"
timeout 1500 python bench.py > gpurun_out/bench_r2u.json 2> gpurun_out/bench_r2u.err
echo "bench exit $?"; tail -c 600 gpurun_out/bench_r2u.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2u_reference.json 2> gpurun_out/bench_r2u_reference.err
echo "reference arm exit $?"; tail -c 300 gpurun_out/bench_r2u_reference.json
