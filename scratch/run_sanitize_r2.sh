#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  echo "== $tool"
  timeout -s KILL 600 compute-sanitizer --tool $tool python scratch/sanitize_r2.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|^ok|Error|hazard|Invalid" | head -8
done 2>&1 | tee gpurun_out/sanitize_r2.txt
