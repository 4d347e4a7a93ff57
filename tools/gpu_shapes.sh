#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out; : > gpurun_out/shapes.jsonl
for shp in "256,2048,365" "1024,1024,1204" "2048,1024,1204" "4096,2048,1000" "16384,2048,1000" "65536,2048,1000" "16384,512,10000" "65536,1024,1204"; do
  steps=2000; [ "${shp%%,*}" -ge 16384 ] && steps=200
  timeout 300 python bench.py --shape $shp --steps $steps --warmup 20 --no-cpu-baseline 2> gpurun_out/shape_err.log | tail -1 > gpurun_out/shape_line.json
  if [ -s gpurun_out/shape_line.json ]; then cat gpurun_out/shape_line.json >> gpurun_out/shapes.jsonl; python - <<'PY'
import json
d=json.loads(open('gpurun_out/shape_line.json').read())
c=d['config']
print('%-18s %8.2f us/step %8.2fM samples/s  step_frac %.3f  launches/step %d | ' % (f"{c['B_per_gpu']}x{c['D']}x{c['C']}", d['ms_per_step']*1e3, d['value']/1e6, d['roofline']['step_frac'], d['gpu_launches']/d['steps']) + '  '.join('%s %.1fus %s %.2f' % (k['kernel'].replace('_bf16',''), k['us'], k['bound'], k['frac']) for k in d['kernels']))
PY
  else echo "$shp FAILED"; tail -3 gpurun_out/shape_err.log; fi
done
