#!/bin/bash
# Round-2 evidence for profiles/: bench lines of every BASELINE config (1 GPU), ncu launch list of the bench command,
# one `ncu --set full` capture of the one-launch step.  (Each ncu run directly after the same command exited 0.)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out; rm -f gpurun_out/shapes_r2.jsonl gpurun_out/summary.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
line() { timeout 400 python bench.py --steps 200 --warmup 20 --no-e2e-alt "$@" > gpurun_out/_line.log 2>gpurun_out/_line.err; rc=$?; echo "bench $* exit $rc" | tee -a gpurun_out/summary.txt
  grep -E "^\{" gpurun_out/_line.log >> gpurun_out/shapes_r2.jsonl; [ $rc -ne 0 ] && tail -5 gpurun_out/_line.err; }
line --shape 256,2048,1000
for v in raw smooth rel normit gombit base2 base10; do line --shape 256,2048,365 --variant $v --no-cpu-baseline --no-torch-baseline; done
line --shape 1024,1024,1204
line --shape 2048,1024,1204 --no-cpu-baseline
line --shape 1024,1024,1204 --loss sigmoid
line --shape 2048,1024,1204 --loss sigmoid
line --shape 128,64,10 --no-torch-baseline
line --shape 16384,2048,1000 --no-cpu-baseline --steps 40 --warmup 5
line --shape 65536,2048,1000 --no-cpu-baseline --no-torch-baseline --steps 20 --warmup 3
line --shape 16384,512,10000 --no-cpu-baseline --no-torch-baseline --steps 20 --warmup 3
python - <<'PY'
import json
for l in open('gpurun_out/shapes_r2.jsonl'):
    d=json.loads(l); c=d["config"]
    print("%-28s %-8s %8.2f us/step %8.2f M/s  e2e %8.2f us  step_frac %.3f  kernels %s" % (f'{c["B_per_gpu"]}x{c["D"]}x{c["C"]}', c.get("loss","")+"/"+c["variant"][:6], d["ms_per_step"]*1e3, d["value"]/1e6, d["e2e"]["ms_per_step"]*1e3, d["roofline"]["step_frac"], [(k["kernel"][:12], round(k["us"],1), round(k["frac"],2)) for k in d["kernels"]]))
PY
CMD="python bench.py --steps 48 --warmup 3 --no-graph --no-cpu-baseline --no-torch-baseline --no-e2e-alt"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list exit $?" | tee -a gpurun_out/summary.txt
timeout 200 python tools/fused_timing.py > gpurun_out/plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:head_fused -s 4 -c 2 -f -o gpurun_out/prof_r2_fused python tools/fused_timing.py > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?" | tee -a gpurun_out/summary.txt
tail -2 gpurun_out/ncu_full.log
cat gpurun_out/summary.txt
