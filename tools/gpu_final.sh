#!/bin/bash
# End-of-round validation on one B200: parity tests, smoke, bench (both arms), phase timeline, ncu evidence.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 300 > gpurun_out/test_gpu_all.log 2>&1; echo "pytest gpu exit $?"; tail -2 gpurun_out/test_gpu_all.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.log 2>&1; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 200 --warmup 5 > gpurun_out/bench_ref.log 2>&1; echo "bench reference exit $?"; tail -1 gpurun_out/bench_ref.log | cut -c1-200
timeout 120 python tools/tc_timing.py > gpurun_out/tc_timing.txt 2>&1; echo "tc_timing exit $?"
bash tools/gpu_profile.sh
