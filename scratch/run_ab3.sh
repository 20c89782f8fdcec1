#!/bin/bash
cd $GRAFT_REPO_ROOT
bash scratch/ab.sh scratch/lib_base.so scratch/lib_fma.so 2
cp scratch/lib_fma.so veloslam_b200/libveloslam_b200.so
timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -q 2>&1 | tail -15
cp scratch/lib_base.so veloslam_b200/libveloslam_b200.so
