#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 300 -x > gpurun_out/test_gpu_all.log 2>&1; echo "pytest gpu exit $?"; tail -3 gpurun_out/test_gpu_all.log
bash tools/gpu_shapes.sh
