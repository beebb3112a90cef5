#!/bin/bash
# GPU job 47: full GPU suite + smoke + default bench on the CTA-pair convolution kernel
mkdir -p gpurun_out
: > gpurun_out/rc.txt
( time timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider ) > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/rc.txt
grep -E "passed|failed" gpurun_out/gpu_tests.log | tail -2; grep -E "^FAILED|^ERROR" gpurun_out/gpu_tests.log | head
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/rc.txt
( time timeout 900 python bench.py ) > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/rc.txt
tail -2 gpurun_out/smoke.log
python tools/bench_summary.py gpurun_out/bench.log 2>/dev/null | head -40
