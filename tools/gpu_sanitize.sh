#!/bin/bash
# one compute-sanitizer tool per gpurun call (B200_PROFILING.md); memcheck over a small parity subset
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 77 --launch-timeout 0 \
  python -m pytest tests/test_gpu_gemm.py tests/test_gpu_modules.py tests/test_gpu_loss.py -q -m gpu -x --timeout 900 \
  -k "130-200-129 or 77-520 or 3-72-5 or fused_into_backward_launch or head_pipeline or (mixup_fused_vs_oracle and 300) or (vs_oracle and 128-10)" \
  > gpurun_out/sanitizer_memcheck.log 2>&1
echo "memcheck exit $?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/sanitizer_memcheck.log | head -20
