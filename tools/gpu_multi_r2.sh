#!/bin/bash
# Round-2 multi-GPU run: all-reduce vs NCCL (values, time, geometry sweep), then bench.py under torchrun with variants.
# usage: gpurun --gpus N -- bash tools/gpu_multi_r2.sh N "<bench flags>" ...
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
N=${1:-2}; shift
mkdir -p gpurun_out
export IIF_B200_PEER_TIMEOUT_S=20
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
[ "$N" = "2" ] && export AR_SWEEP=1
if [ "$SKIP_CHECK" != "1" ]; then
timeout 300 $RUN tools/check_allreduce.py > gpurun_out/check_allreduce_n$N.log 2>&1; echo "check_allreduce exit $?"
grep -vE "^\*|OMP_NUM" gpurun_out/check_allreduce_n$N.log | tail -40
fi
i=0
for v in "$@"; do
  i=$((i+1))
  echo "== N=$N bench $v"
  algo=auto
  fence=gpu
  case "$v" in push:*) algo=push; v="${v#push:}";; esac
  case "$v" in pull:*) algo=pull; v="${v#pull:}";; esac
  np=$N
  case "$v" in n4:*) np=4; v="${v#n4:}";; esac
  case "$v" in n2:*) np=2; v="${v#n2:}";; esac
  RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29511"
  case "$v" in sysfence:*) fence=sys; v="${v#sysfence:}";; esac
  IIF_B200_AR_MIDFENCE=$fence IIF_B200_AR_ALGO=$algo timeout 400 $RUN bench.py --gpus $np $v > gpurun_out/bench_n${N}_$i.log 2>&1
  echo "exit $?"
  grep -E "^\{" gpurun_out/bench_n${N}_$i.log | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l)
    print('value %.2fM  %.2f us/step (region min %.3f max %.3f ms, %d repeats) | e2e %.2fM %.2f us/step | n_gpus %d' % (d['value']/1e6, d['ms_per_step']*1e3, d['region_ms']['min'], d['region_ms']['max'], d['repeats'], d['e2e']['value']/1e6, d['e2e']['ms_per_step']*1e3, d['n_gpus']))
    print(d['config']['parallelism']); print('allreduce_check', d['allreduce_check'])"
  grep -iE "error|timed out|Traceback" gpurun_out/bench_n${N}_$i.log | head -5
done
