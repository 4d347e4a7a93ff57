#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 120 python tools/tc_timing.py > gpurun_out/tc_timing.txt 2>&1; echo "exit $?"
sed -n '/whole step/,$p' gpurun_out/tc_timing.txt
