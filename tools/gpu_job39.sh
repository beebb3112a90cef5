#!/bin/bash
# GPU job 39: full GPU suite on the tensor-core dense layer + the row-run trim of camera_project; kernel timings; short bench
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider ) > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?" > gpurun_out/rc.txt
grep -E "passed|failed" gpurun_out/gpu_tests.log | tail -2; grep -E "^FAILED|^ERROR" gpurun_out/gpu_tests.log | head
timeout 300 python tests/perf_kernels.py camera > gpurun_out/perf_camera.log 2>&1; echo "perf rc=$?" >> gpurun_out/rc.txt
tail -4 gpurun_out/perf_camera.log
timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-configs --no-alt > gpurun_out/bench_short.log 2> gpurun_out/bench_short.err; echo "bench rc=$?" >> gpurun_out/rc.txt
python tools/bench_summary.py gpurun_out/bench_short.log 2>/dev/null | head -12
cat gpurun_out/rc.txt
