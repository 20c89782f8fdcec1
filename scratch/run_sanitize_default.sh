#!/bin/bash
cd $GRAFT_REPO_ROOT
for tool in memcheck racecheck synccheck; do
  echo "== $tool"
  timeout -s KILL 300 compute-sanitizer --tool $tool python scratch/sanitize.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|^ok|Error|hazard" | head -5
done
