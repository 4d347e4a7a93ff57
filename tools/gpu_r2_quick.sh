#!/bin/bash
# Quick single-GPU re-check after a library change: every GPU test, the fp32 modes, the driver's bench command.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests -q -m gpu --timeout 300 > gpurun_out/test_gpu_all.log 2>&1; echo "pytest gpu exit $?" | tee -a gpurun_out/summary.txt
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/test_gpu_all.log | head -30
timeout 300 python tools/fp32_modes.py > gpurun_out/fp32_modes.txt 2>&1; echo "fp32 modes exit $?" | tee -a gpurun_out/summary.txt; cat gpurun_out/fp32_modes.txt
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench.log 2>gpurun_out/bench.err; echo "bench exit $?" | tee -a gpurun_out/summary.txt
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
    print("bench: us/step %.2f value %.3fM e2e %.2f us modes %s" % (d["ms_per_step"] * 1e3, d["value"] / 1e6, d["e2e"]["ms_per_step"] * 1e3, d["e2e"]["modes"]))
    print("e2e bound", d["e2e"].get("bound"))
except Exception as e:
    print("bench parse failed", e); print(open('gpurun_out/bench.log').read()[-2000:]); print(open('gpurun_out/bench.err').read()[-2000:])
PY
for s in 512,2048,1000 1024,1024,1204; do timeout 300 python bench.py --shape $s --steps 200 --warmup 20 --no-cpu-baseline --no-torch-baseline --no-e2e-alt 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$s: %.2f us/step, %d launches/step, e2e %.1f us, bound %s' % (d['ms_per_step'] * 1e3, d['gpu_launches'] // d['steps'], d['e2e']['ms_per_step'] * 1e3, d['e2e'].get('bound')))"; done
cat gpurun_out/summary.txt
