#!/bin/bash
# Bench lines (1 GPU) of every BASELINE config with the final library: profiles/r2_bench_shapes.jsonl
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out; rm -f gpurun_out/shapes_r2.jsonl gpurun_out/summary.txt
line() { timeout 400 python bench.py --steps 200 --warmup 20 --no-e2e-alt "$@" > gpurun_out/_line.log 2>gpurun_out/_line.err; rc=$?; echo "bench $* exit $rc" | tee -a gpurun_out/summary.txt
  grep -E "^\{" gpurun_out/_line.log >> gpurun_out/shapes_r2.jsonl; [ $rc -ne 0 ] && tail -5 gpurun_out/_line.err; }
line --shape 256,2048,1000 --cpu-seconds 4
for v in raw smooth rel normit gombit base2 base10; do line --shape 256,2048,365 --variant $v --no-cpu-baseline --no-torch-baseline; done
line --shape 1024,1024,1204 --no-cpu-baseline
line --shape 2048,1024,1204 --no-cpu-baseline
line --shape 1024,1024,1204 --loss sigmoid --no-cpu-baseline
line --shape 2048,1024,1204 --loss sigmoid --no-cpu-baseline
line --shape 128,64,10 --no-torch-baseline --no-cpu-baseline
line --shape 16384,2048,1000 --no-cpu-baseline --steps 40 --warmup 5
line --shape 65536,2048,1000 --no-cpu-baseline --no-torch-baseline --steps 20 --warmup 3
line --shape 16384,512,10000 --no-cpu-baseline --no-torch-baseline --steps 20 --warmup 3
python - <<'PY'
import json
for l in open('gpurun_out/shapes_r2.jsonl'):
    d=json.loads(l); c=d["config"]
    print("%-28s %-8s %8.2f us/step %8.2f M/s  e2e %8.2f us  step_frac %.3f  kernels %s" % (f'{c["B_per_gpu"]}x{c["D"]}x{c["C"]}', c.get("loss","")+"/"+c["variant"][:6], d["ms_per_step"]*1e3, d["value"]/1e6, d["e2e"]["ms_per_step"]*1e3, d["roofline"]["step_frac"], [(k["kernel"][:12], round(k["us"],1), round(k["frac"],2)) for k in d["kernels"]]))
PY
