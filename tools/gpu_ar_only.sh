#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
N=${1:-2}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/check_allreduce.py > gpurun_out/check_allreduce_n$N.log 2>&1
echo "check exit $?"; grep -vE "^W|^\[W|Warning|^\*|OMP_NUM|^$|buf [123]" gpurun_out/check_allreduce_n$N.log | tail -24
