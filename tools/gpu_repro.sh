#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
for v in "--no-prefetch" "--no-fused-loss" "--no-prefetch --no-fused-loss" ""; do
  echo "== eager $v"
  timeout 120 python bench.py --steps 48 --warmup 3 --no-graph --no-cpu-baseline $v > gpurun_out/repro.log 2>&1
  echo "exit $?"; grep -m3 "timed out" gpurun_out/repro.log; tail -1 gpurun_out/repro.log | cut -c1-200
done
