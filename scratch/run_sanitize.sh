#!/bin/bash
# compute-sanitizer over the small decode workloads, both pipelines
cd $GRAFT_REPO_ROOT
for sp in 0 1; do
  for tool in memcheck racecheck synccheck; do
    echo "== single_pass=$sp $tool"
    VELOSLAM_SINGLE_PASS=$sp timeout -s KILL 400 compute-sanitizer --tool $tool python scratch/sanitize.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|^ok|Error|hazard" | head -8
  done
done
