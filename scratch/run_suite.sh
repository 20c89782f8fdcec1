#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout -s KILL 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout -s KILL 200 python __graft_entry__.py smoke 2>&1 | tail -1
