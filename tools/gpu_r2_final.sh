#!/bin/bash
# Final single-GPU validation of round 2: every GPU test, smoke, the driver's bench command (both arms), the routing
# boundary between the one-launch step and the multi-launch chain, the tiny shape, the fp32 modes.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt gpurun_out/route_r2.jsonl
timeout 1200 python -m pytest tests -q -m gpu --timeout 300 > gpurun_out/test_gpu_all.log 2>&1; echo "pytest gpu exit $?" | tee -a gpurun_out/summary.txt
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/test_gpu_all.log | head -30
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/summary.txt; tail -3 gpurun_out/smoke.log
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ref.log 2>&1; echo "bench reference exit $?" | tee -a gpurun_out/summary.txt; tail -1 gpurun_out/bench_ref.log | cut -c1-400
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench.log 2>gpurun_out/bench.err; echo "bench exit $?" | tee -a gpurun_out/summary.txt
tail -1 gpurun_out/bench.log | cut -c1-1500
line() { tag="$1"; shift; timeout 300 python bench.py --steps 200 --warmup 20 --no-e2e-alt --no-cpu-baseline --no-torch-baseline "$@" > gpurun_out/_line.log 2>gpurun_out/_line.err; rc=$?
  echo "bench [$tag] $* exit $rc" | tee -a gpurun_out/summary.txt
  grep -E "^\{" gpurun_out/_line.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); d['route'] = '$tag'; print(json.dumps(d))
    print('   %s %s: %.2f us/step, %d launches/step' % ('$tag', d['config']['workload'][:40], d['ms_per_step'] * 1e3, d['gpu_launches'] // max(d['steps'], 1)), file=sys.stderr)
" >> gpurun_out/route_r2.jsonl; [ $rc -ne 0 ] && tail -5 gpurun_out/_line.err; }
for s in 512,2048,1000 1024,2048,1000 512,1024,1204 1024,1024,1204 2048,1024,1204 2048,2048,1000; do
  IIF_B200_FUSED_MAX_ROW_PASSES=0 line one-launch --shape $s
  line chain --shape $s --no-persistent
  line default --shape $s
done
line tiny --shape 128,64,10
timeout 300 python tools/fp32_modes.py > gpurun_out/fp32_modes.txt 2>&1; echo "fp32 modes exit $?" | tee -a gpurun_out/summary.txt; cat gpurun_out/fp32_modes.txt
cat gpurun_out/summary.txt
