#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?" > gpurun_out/rc.txt
grep -E "passed|failed" gpurun_out/gpu_tests.log | tail -2; grep -E "^FAILED|^ERROR" gpurun_out/gpu_tests.log | head
for i in 1 2; do
timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-alt --no-configs --steps 30 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/rc.txt
python tools/bench_summary.py gpurun_out/bench.log 2>/dev/null | grep -E "^value" | cut -c1-100
done
cat gpurun_out/rc.txt
