#!/bin/bash
# GPU job 22: border_expand with 16-byte stores, plain bilinear resize on 32-bit index arithmetic: parity + bench
mkdir -p gpurun_out
: > gpurun_out/rc.txt
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/rc.txt
grep -E "passed|failed" gpurun_out/gpu_tests.log | tail -2
grep -E "^FAILED|^ERROR" gpurun_out/gpu_tests.log | head
( time timeout 900 python bench.py --no-cpu-baseline --no-configs ) > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/rc.txt
python tools/bench_summary.py gpurun_out/bench.log 2>/dev/null | head -8
cat gpurun_out/rc.txt
