#!/bin/bash
# N-GPU bench (weak scaling, NCCL all-reduce of dW+db).  Usage: gpurun --gpus N -- bash tools/gpu_multi.sh N
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
for v in "" "--sync-allreduce"; do
  echo "== N=$N $v"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 3000 --warmup 50 $v > gpurun_out/bench_n$N.log 2>&1
  echo "exit $?"
  grep -E "^\{" gpurun_out/bench_n$N.log | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l)
    print('value %.2fM  %.2f us/step | e2e %.2fM %.2f us/step | n_gpus %d' % (d['value']/1e6, d['ms_per_step']*1e3, d['e2e']['value']/1e6, d['e2e']['ms_per_step']*1e3, d['n_gpus']))
    print(d['config']['parallelism'])"
  grep -iE "error|timed out|Traceback" gpurun_out/bench_n$N.log | head -5
done
echo "== N=1 reference"; timeout 300 python bench.py --steps 3000 --warmup 50 --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.2fM  %.2f us/step | e2e %.2fM' % (d['value']/1e6, d['ms_per_step']*1e3, d['e2e']['value']/1e6))"
