#!/bin/bash
# bench variants for A/B comparison
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
for v in "" "--no-prefetch" "--no-fused-loss" "--no-fused-loss --no-prefetch" ""; do
  echo "== bench.py $v"
  timeout 300 python bench.py --no-cpu-baseline $v 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.2fM  %.2f us/step | e2e %.2fM %.2f us/step | launches/step %d' % (d['value']/1e6, d['ms_per_step']*1e3, d['e2e']['value']/1e6, d['e2e']['ms_per_step']*1e3, d['gpu_launches']/d['steps']))
for k in d['kernels']: print('   ', k['kernel'], round(k['us'],2))"
done
