#!/bin/bash
# ncu --set full of the small-batch kernels (rotation-sized batches, direct launches)
sed -i "s/n_rot = 300/n_rot = 30/" scratch/online_probe.py
for k in k_pose k_frames k_scan; do
VELOSLAM_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:^$k -s 10 -c 1 -o gpurun_out/r2k_online_$k -f python scratch/online_probe.py > gpurun_out/r2k_online_$k.log 2>&1
done
ls -la gpurun_out/r2k_online_*
