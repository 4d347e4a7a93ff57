#!/bin/bash
# first GPU shakedown: each test file in its own process (a device trap must not cascade)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for f in test_gpu_loss test_gpu_hist test_gpu_gemm test_gpu_modules; do
  timeout 600 python -m pytest tests/$f.py -q -m gpu -x --timeout 300 > gpurun_out/$f.log 2>&1
  echo "$f exit $?" >> gpurun_out/summary.txt
  tail -5 gpurun_out/$f.log
done
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py > gpurun_out/bench.log 2>&1; echo "bench exit $?" >> gpurun_out/summary.txt
tail -3 gpurun_out/bench.log
cat gpurun_out/summary.txt
