#!/bin/bash
# final-HEAD evidence: launch list of the bench command (headline leg only) + ncu --set full of k_decode
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
LEAN="--no-e2e --no-cpu --no-online --no-deskew --no-single-pass --no-parity --no-facade --no-hdl32 --recording-leg-hours 0 --online-udp-seconds 0"
timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/r2o_launches.csv python bench.py --steps 2 --warmup 1 $LEAN > gpurun_out/r2o_launches_bench.log 2>&1
echo "launch list exit $?"; grep -c k_decode gpurun_out/r2o_launches.csv
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:"^k_decode\$" -s 2 -c 1 -f \
  -o gpurun_out/r2o_k_decode python scratch/prof_step.py > gpurun_out/r2o_ncu_k_decode.log 2>&1
echo "ncu exit $?"; ls -la gpurun_out/r2o_k_decode.ncu-rep
python bench.py --steps 10 --warmup 3 $LEAN
