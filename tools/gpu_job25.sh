#!/bin/bash
# GPU job 25: independent branches of the step on side streams (parallel branches of the CUDA graph): parity, A/B bench
mkdir -p gpurun_out
: > gpurun_out/rc.txt
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/rc.txt
grep -E "passed|failed" gpurun_out/gpu_tests.log | tail -2
grep -E "^FAILED|^ERROR" gpurun_out/gpu_tests.log | head
for mode in 0 1 0 1; do
  B200BEV_BENCH_SERIAL=$mode timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-alt --steps 30 > gpurun_out/bench_ab_$mode.log 2> gpurun_out/bench_ab.err; echo "bench serial=$mode rc=$?" >> gpurun_out/rc.txt
  echo "serial=$mode"; python tools/bench_summary.py gpurun_out/bench_ab_$mode.log 2>/dev/null | grep -E "^value|lidar_only|camera_only|fusion  |full_split|stress" | cut -c1-90
done
cat gpurun_out/rc.txt
