#!/bin/bash
# Round-2 inner loop: fused-step parity, timeline of the head shapes, a quick bench; `ncu` adds the source-level capture.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_fused.py -q -m gpu --timeout 120 -x > gpurun_out/test_gpu_fused.log 2>&1
echo "test_gpu_fused exit $?" | tee -a gpurun_out/summary.txt
tail -5 gpurun_out/test_gpu_fused.log
for shp in 256,2048,1000 256,2048,365 1024,1024,1204; do
  timeout 200 python tools/fused_timing.py $shp > gpurun_out/fused_timing_$shp.txt 2>&1; echo "fused_timing $shp exit $?" | tee -a gpurun_out/summary.txt
  grep -A28 "L2 flushed" gpurun_out/fused_timing_$shp.txt; tail -3 gpurun_out/fused_timing_$shp.txt
done
for extra in "" "--shape 256,2048,365" "--shape 1024,1024,1204" "--shape 2048,1024,1204"; do
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-torch-baseline --no-e2e-alt $extra > gpurun_out/bench_quick.log 2>&1; echo "bench $extra exit $?" | tee -a gpurun_out/summary.txt
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_quick.log').read().strip().splitlines()[-1])
    print("bench %s: us/step %.2f value %.3fM e2e %.2f us step_frac %.3f" % (d["config"]["workload"][:40], d["ms_per_step"]*1e3, d["value"]/1e6, d["e2e"]["ms_per_step"]*1e3, d["roofline"]["step_frac"]))
except Exception as e:
    print("bench parse failed", e); print(open('gpurun_out/bench_quick.log').read()[-2000:])
PY
done
if [ "$1" = "ncu" ]; then
timeout 200 python tools/fused_timing.py > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:head_fused -s 4 -c 2 -f -o gpurun_out/prof_r2_fused python tools/fused_timing.py > gpurun_out/ncu_full.log 2>&1
echo "ncu exit $?" | tee -a gpurun_out/summary.txt
tail -3 gpurun_out/ncu_full.log
fi
cat gpurun_out/summary.txt
