#!/bin/bash
cd $GRAFT_REPO_ROOT
cp veloslam_b200/libveloslam_b200.so /tmp/orig.so
for v in new nomath nostore neither; do
  cp scratch/lib_$v.so veloslam_b200/libveloslam_b200.so
  echo "== $v"
  timeout -s KILL 100 python scratch/prof_fused.py 2>&1 | grep decode_ms
done
cp /tmp/orig.so veloslam_b200/libveloslam_b200.so
