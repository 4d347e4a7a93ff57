#!/bin/bash
# One short validation of a kernel change: every GPU test, then the driver's bench command and the phase timeline.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 200 python -m pytest tests -q -x -m gpu --timeout 60 > gpurun_out/test_gpu_all.log 2>&1; rc=$?; echo "pytest gpu exit $rc"
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/test_gpu_all.log | head -10
[ $rc -ne 0 ] && { tail -30 gpurun_out/test_gpu_all.log; exit 1; }
timeout 120 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-torch-baseline > gpurun_out/bench_oneshot.log 2>&1; echo "bench exit $?"
tail -1 gpurun_out/bench_oneshot.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('us/step %.2f  e2e %.2f us  modes %s' % (d['ms_per_step'] * 1e3, d['e2e']['ms_per_step'] * 1e3, d['e2e']['modes']))"
timeout 60 python tools/fused_timing.py > gpurun_out/fused_timing.txt 2>&1; echo "timing exit $?"
grep -A24 "L2 flushed" gpurun_out/fused_timing.txt | head -26; tail -3 gpurun_out/fused_timing.txt
