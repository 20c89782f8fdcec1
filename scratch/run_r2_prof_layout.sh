#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
TAG=${1:-r2b}
PROF_STEPS=1 timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:'^k_layout$' -c 1 -f \
    -o gpurun_out/prof_k_layout_$TAG python scratch/prof_step.py > gpurun_out/ncu_k_layout_$TAG.log 2>&1
tail -3 gpurun_out/ncu_k_layout_$TAG.log | cut -c1-200
