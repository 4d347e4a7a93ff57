#!/bin/bash
# Round-2 inner loop: fused-step parity, timeline variants (env-forced plans), optional ncu source-level capture.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_fused.py -q -m gpu --timeout 120 -x > gpurun_out/test_gpu_fused.log 2>&1
echo "test_gpu_fused exit $?" | tee -a gpurun_out/summary.txt
tail -5 gpurun_out/test_gpu_fused.log
timeout 200 python tools/fused_timing.py > gpurun_out/fused_timing.txt 2>&1; echo "fused_timing exit $?" | tee -a gpurun_out/summary.txt
tail -60 gpurun_out/fused_timing.txt
IIF_B200_FUSED_SDX=4 timeout 200 python tools/fused_timing.py > gpurun_out/fused_timing_sdx4.txt 2>&1; echo "fused_timing(SDX=4) exit $?" | tee -a gpurun_out/summary.txt
grep -A32 "L2 flushed" gpurun_out/fused_timing_sdx4.txt; tail -3 gpurun_out/fused_timing_sdx4.txt
IIF_B200_COOP=0 timeout 200 python tools/fused_timing.py > gpurun_out/fused_timing_nocoop.txt 2>&1; echo "fused_timing(COOP=0) exit $?" | tee -a gpurun_out/summary.txt
tail -3 gpurun_out/fused_timing_nocoop.txt
for shp in 256,2048,365 1024,1024,1204; do
  timeout 200 python tools/fused_timing.py $shp > gpurun_out/fused_timing_$shp.txt 2>&1; echo "fused_timing $shp exit $?" | tee -a gpurun_out/summary.txt
  grep -A32 "L2 flushed" gpurun_out/fused_timing_$shp.txt; tail -3 gpurun_out/fused_timing_$shp.txt
done
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-torch-baseline > gpurun_out/bench_quick.log 2>&1; echo "bench exit $?" | tee -a gpurun_out/summary.txt
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_quick.log').read().strip().splitlines()[-1])
    print("bench: us/step", d["ms_per_step"]*1e3, "value", d["value"], "e2e", d["e2e"]["ms_per_step"]*1e3, d["e2e"]["modes"], "launch", d["config"]["launch"])
except Exception as e:
    print("bench parse failed", e); print(open('gpurun_out/bench_quick.log').read()[-2000:])
PY
IIF_B200_COOP=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-torch-baseline > gpurun_out/bench_quick_nocoop.log 2>&1; echo "bench(COOP=0) exit $?" | tee -a gpurun_out/summary.txt
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_quick_nocoop.log').read().strip().splitlines()[-1])
    print("bench COOP=0: us/step", d["ms_per_step"]*1e3, "value", d["value"], "e2e", d["e2e"]["ms_per_step"]*1e3, d["e2e"]["modes"])
except Exception as e:
    print("bench parse failed", e); print(open('gpurun_out/bench_quick_nocoop.log').read()[-2000:])
PY
if [ "$1" = "ncu" ]; then
timeout 200 python tools/fused_timing.py > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:head_fused -s 4 -c 2 -f -o gpurun_out/prof_r2_fused python tools/fused_timing.py > gpurun_out/ncu_full.log 2>&1
echo "ncu exit $?" | tee -a gpurun_out/summary.txt
tail -3 gpurun_out/ncu_full.log
fi
cat gpurun_out/summary.txt
