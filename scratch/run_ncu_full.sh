#!/bin/bash
# one ncu --set full capture (with source) of k_decode in the bench batch
cd $GRAFT_REPO_ROOT
timeout -s KILL 500 ncu --set full --clock-control none --import-source on -k regex:k_decode -s 4 -c 1 -f -o gpurun_out/prof_r1z python bench.py --no-cpu --no-e2e --no-online --no-deskew --steps 2 --warmup 3 > gpurun_out/ncu_full_r1z.log 2>&1
tail -3 gpurun_out/ncu_full_r1z.log | cut -c1-300
ls -la gpurun_out/prof_r1z.ncu-rep
