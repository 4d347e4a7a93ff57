#!/bin/bash
# Round-2 first contact: the one-launch step's parity tests, the full GPU suite, timeline, bench.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_fused.py -q -m gpu --timeout 120 -x > gpurun_out/test_gpu_fused.log 2>&1
echo "test_gpu_fused exit $?" | tee -a gpurun_out/summary.txt
tail -25 gpurun_out/test_gpu_fused.log
timeout 200 python tools/fused_timing.py > gpurun_out/fused_timing.txt 2>&1; echo "fused_timing exit $?" | tee -a gpurun_out/summary.txt
cat gpurun_out/fused_timing.txt | tail -45
timeout 900 python -m pytest tests -q -m gpu --timeout 300 --deselect tests/test_gpu_fused.py > gpurun_out/test_gpu_all.log 2>&1
echo "pytest gpu (rest) exit $?" | tee -a gpurun_out/summary.txt
tail -8 gpurun_out/test_gpu_all.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>gpurun_out/bench.err; echo "bench exit $?" | tee -a gpurun_out/summary.txt
tail -c 6000 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-persistent --no-cpu-baseline --no-torch-baseline > gpurun_out/bench_r1path.log 2>&1; echo "bench(no-persistent) exit $?" | tee -a gpurun_out/summary.txt
tail -c 1500 gpurun_out/bench_r1path.log
cat gpurun_out/summary.txt
