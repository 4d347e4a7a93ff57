#!/bin/bash
# One 8-GPU box: all-reduce vs NCCL at N=8, then bench.py at N = 8, 4, 2, 1 (driver config and a long run).
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
export IIF_B200_PEER_TIMEOUT_S=20
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
run() { n=$1; shift; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 "$@"; }
run 8 tools/check_allreduce.py > gpurun_out/check_allreduce_n8.log 2>&1; echo "check_allreduce n8 exit $?"
grep -vE "^\*|OMP_NUM" gpurun_out/check_allreduce_n8.log | tail -14
show() { grep -E "^\{" $1 | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l)
    print('N=%d value %.2fM  %.2f us/step (region min %.3f max %.3f ms, %d repeats) | e2e %.2fM %.2f us/step' % (d['n_gpus'], d['value']/1e6, d['ms_per_step']*1e3, d['region_ms']['min'], d['region_ms']['max'], d['repeats'], d['e2e']['value']/1e6, d['e2e']['ms_per_step']*1e3))
    print('   ', d['config']['parallelism'][:150]); print('    allreduce_check', d['allreduce_check'])"; grep -iE "error|timed out|Traceback" $1 | head -5; }
NS="8 4 2"; [ "$1" = "quick" ] && NS="8 4"
for N in $NS; do
  for v in "--steps 20 --warmup 5 --no-e2e-alt" "--steps 2000 --warmup 50 --no-e2e-alt"; do
    [ "$1" = "quick" ] && [ "$N" != "8" ] && [ "$v" != "--steps 20 --warmup 5 --no-e2e-alt" ] && continue
    tag=$(echo $v | tr -d ' -')
    echo "== N=$N $v"
    run $N bench.py --gpus $N $v > gpurun_out/scale_n${N}_$tag.log 2>&1; echo "exit $?"
    show gpurun_out/scale_n${N}_$tag.log
  done
done
[ "$1" = "quick" ] && exit 0
echo "== N=8 NCCL arm"; run 8 bench.py --gpus 8 --steps 2000 --warmup 50 --no-e2e-alt --allreduce nccl > gpurun_out/scale_n8_nccl.log 2>&1; echo "exit $?"; show gpurun_out/scale_n8_nccl.log
echo "== N=8 push form"; IIF_B200_AR_ALGO=push run 8 bench.py --gpus 8 --steps 2000 --warmup 50 --no-e2e-alt > gpurun_out/scale_n8_push.log 2>&1; echo "exit $?"; show gpurun_out/scale_n8_push.log
echo "== N=1"; timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-torch-baseline > gpurun_out/scale_n1.log 2>&1; echo "exit $?"
grep -E "^\{" gpurun_out/scale_n1.log | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('N=1 value %.2fM %.2f us/step e2e %.2f us' % (d['value']/1e6, d['ms_per_step']*1e3, d['e2e']['ms_per_step']*1e3))"
