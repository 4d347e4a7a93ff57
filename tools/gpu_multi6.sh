#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
N=${1:-2}; shift
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_allreduce.py -q -m gpu -x 2>&1 | tail -2
for v in "$@"; do
  echo "== N=$N $v"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 3000 --warmup 50 --allreduce peer $v > gpurun_out/bench_n${N}.log 2>&1
  echo "exit $?"
  grep -E "^\{" gpurun_out/bench_n$N.log | tee gpurun_out/bench_n${N}_$(echo $v | tr -d ' -').json | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l)
    print('value %.2fM  %.2f us/step | e2e %.2fM %.2f us/step | n_gpus %d' % (d['value']/1e6, d['ms_per_step']*1e3, d['e2e']['value']/1e6, d['e2e']['ms_per_step']*1e3, d['n_gpus']))
    print(d['config']['parallelism'])"
  grep -iE "error|timed out|Traceback" gpurun_out/bench_n$N.log | head -5
done
