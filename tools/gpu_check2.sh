#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 300 -x > gpurun_out/test_gpu_all.log 2>&1; echo "pytest gpu exit $?"; tail -3 gpurun_out/test_gpu_all.log
echo "== eager stress"; timeout 300 python bench.py --steps 3000 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/eager.log 2>&1; echo "exit $?"; grep -m2 "timed out" gpurun_out/eager.log; tail -1 gpurun_out/eager.log | cut -c1-230
bash tools/gpu_profile.sh
