#!/bin/bash
# GPU check: parity tests, smoke, bench, per-CTA phase timeline.  Usage: gpurun -- bash tools/gpu_check.sh [quick]
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
for f in test_gpu_gemm test_gpu_loss test_gpu_modules test_gpu_hist; do
  timeout 900 python -m pytest tests/$f.py -q -m gpu --timeout 300 -x > gpurun_out/$f.log 2>&1
  echo "$f exit $?" >> gpurun_out/summary.txt
  grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/$f.log | head -20
done
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py > gpurun_out/bench.log 2>&1; echo "bench exit $?" >> gpurun_out/summary.txt
tail -1 gpurun_out/bench.log
timeout 120 python tools/tc_timing.py > gpurun_out/tc_timing.txt 2>&1; echo "tc_timing exit $?" >> gpurun_out/summary.txt
cat gpurun_out/tc_timing.txt
cat gpurun_out/summary.txt
