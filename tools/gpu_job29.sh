#!/bin/bash
# GPU job 29: fp32-accuracy MLP with 2 M points per pass (32 x 35,000 in one pass): parity + timing + f32 bench
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?" > gpurun_out/rc.txt
grep -E "passed|failed" gpurun_out/gpu_tests.log | tail -2; grep -E "^FAILED|^ERROR" gpurun_out/gpu_tests.log | head
timeout 300 python tests/probes/split_probe.py 2>&1 | tail -5
timeout 600 python bench.py --precision f32 --no-cpu-baseline --no-e2e --no-configs --no-alt --steps 10 > gpurun_out/bench_f32.log 2> gpurun_out/bench_f32.err; echo "bench rc=$?" >> gpurun_out/rc.txt
python tools/bench_summary.py gpurun_out/bench_f32.log 2>/dev/null | head -3
cat gpurun_out/rc.txt
