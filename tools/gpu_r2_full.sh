#!/bin/bash
# Full single-GPU validation: every GPU test, smoke, timeline, bench (both arms).
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests -q -m gpu --timeout 300 > gpurun_out/test_gpu_all.log 2>&1; echo "pytest gpu exit $?" | tee -a gpurun_out/summary.txt
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/test_gpu_all.log | head -30
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/summary.txt; tail -3 gpurun_out/smoke.log
timeout 200 python tools/fused_timing.py > gpurun_out/fused_timing.txt 2>&1; echo "fused_timing exit $?" | tee -a gpurun_out/summary.txt
grep -A28 "L2 flushed" gpurun_out/fused_timing.txt; tail -3 gpurun_out/fused_timing.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>gpurun_out/bench.err; echo "bench exit $?" | tee -a gpurun_out/summary.txt
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
    print("bench: us/step %.2f value %.3fM e2e %.2f us modes %s" % (d["ms_per_step"]*1e3, d["value"]/1e6, d["e2e"]["ms_per_step"]*1e3, d["e2e"]["modes"]))
    print("roofline", {k: d["roofline"][k] for k in ("kernel","us_per_launch","frac","step_frac")})
    print("cpu", d["cpu_baseline"]["value"], "torch", {k: d["torch_gpu_baseline"][k] for k in ("bf16","fp32")})
except Exception as e:
    print("bench parse failed", e); print(open('gpurun_out/bench.log').read()[-2000:]); print(open('gpurun_out/bench.err').read()[-2000:])
PY
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref.log 2>&1; echo "bench reference exit $?" | tee -a gpurun_out/summary.txt; tail -1 gpurun_out/bench_ref.log | cut -c1-300
cat gpurun_out/summary.txt
