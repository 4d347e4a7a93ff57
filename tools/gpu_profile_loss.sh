#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
CMD="python bench.py --shape 16384,2048,1000 --steps 20 --warmup 3 --no-graph --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/plain_loss.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:row_softmax_kernel -s 10 -c 2 -f -o gpurun_out/prof_r1_loss_16k $CMD > gpurun_out/ncu_loss.log 2>&1
echo "exit $?"; tail -2 gpurun_out/ncu_loss.log
