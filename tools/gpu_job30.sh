#!/bin/bash
# GPU job 30: 3x3 convolution kernel with flat tiling (tiles of 256 flat indices, not whole image rows): parity, conv timings, bench
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?" > gpurun_out/rc.txt
grep -E "passed|failed" gpurun_out/gpu_tests.log | tail -2; grep -E "^FAILED|^ERROR" gpurun_out/gpu_tests.log | head
timeout 300 python tests/perf_kernels.py conv 2>&1 | grep -E "^conv" | cut -c1-110
timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-configs > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/rc.txt
python tools/bench_summary.py gpurun_out/bench.log 2>/dev/null | grep -E "^value|fp32_path" | cut -c1-200
cat gpurun_out/rc.txt
