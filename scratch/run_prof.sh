#!/bin/bash
cd $GRAFT_REPO_ROOT
cp veloslam_b200/libveloslam_b200.so /tmp/orig.so
cp scratch/libprof.so veloslam_b200/libveloslam_b200.so
timeout -s KILL 200 python scratch/${1:-prof_fused.py} 2>&1 | tail -20
cp /tmp/orig.so veloslam_b200/libveloslam_b200.so
