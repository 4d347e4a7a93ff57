#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests -q -m gpu --timeout 300 -x > gpurun_out/test_gpu_all.log 2>&1
echo "tests exit $?" >> gpurun_out/summary.txt; tail -3 gpurun_out/test_gpu_all.log
timeout 300 python bench.py --steps 48 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:gemm_tc_kernel|row_softmax|colsum' -s 20 -c 10 \
  -o gpurun_out/prof_r1_baseline -f python bench.py --steps 48 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
echo "ncu exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
