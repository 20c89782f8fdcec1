#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -x -q 2>&1 | tail -3
bash scratch/ab.sh scratch/lib_base.so scratch/lib_new.so 2
cp scratch/lib_new.so veloslam_b200/libveloslam_b200.so
