#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 4 --steps 2000 --warmup 50 > gpurun_out/bench_n4.log 2>&1
echo "exit $?"
grep -E "^\{" gpurun_out/bench_n4.log | tee gpurun_out/bench_n4.json | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l)
    print('value %.2fM  %.2f us/step | e2e %.2fM %.2f us/step | n_gpus %d' % (d['value']/1e6, d['ms_per_step']*1e3, d['e2e']['value']/1e6, d['e2e']['ms_per_step']*1e3, d['n_gpus']))
    print(d['config']['parallelism'])"
grep -iE "error|timed out|Traceback" gpurun_out/bench_n4.log | head -5
