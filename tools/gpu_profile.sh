#!/bin/bash
# ncu evidence for profiles/: launch list of the bench command + one --set full capture of the head kernels.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
CMD="python bench.py --steps 48 --warmup 3 --no-graph --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list exit $?"
timeout 300 $CMD > gpurun_out/plain2.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 24 -c 4 -f -o gpurun_out/prof_r1_head $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?"
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out/
