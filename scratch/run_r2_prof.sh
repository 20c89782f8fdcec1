#!/bin/bash
# round 2: launch list + one ncu --set full capture per kernel of the path
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
TAG=${1:-r2a}
if [ -z "$SKIP_LIST" ]; then
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv \
  --log-file gpurun_out/launches_$TAG.csv python scratch/prof_step.py > gpurun_out/prof_step_$TAG.log 2>&1
fi
for k in k_scan k_pose k_decode k_layout; do
  timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:"^${k}\$" -s 2 -c 1 -f \
    -o gpurun_out/prof_${k}_$TAG python scratch/prof_step.py > gpurun_out/ncu_${k}_$TAG.log 2>&1
  tail -2 gpurun_out/ncu_${k}_$TAG.log | cut -c1-200
done
ls -la gpurun_out/*_$TAG.ncu-rep
python scratch/prof_step.py
