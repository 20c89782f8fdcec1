#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:"^k_pose\$" -s 2 -c 1 -f \
  -o gpurun_out/r2w_k_pose python scratch/prof_step.py > gpurun_out/r2w_ncu_k_pose.log 2>&1
echo "ncu exit $?"; ls -la gpurun_out/r2w_k_pose.ncu-rep
