#!/bin/bash
# GPU job 35: fp32 (layer-by-layer) fusion path with the lidar / radar branches on side streams: parity, A/B
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?" > gpurun_out/rc.txt
grep -E "passed|failed" gpurun_out/gpu_tests.log | tail -2; grep -E "^FAILED|^ERROR" gpurun_out/gpu_tests.log | head
for mode in 0 1 0 1; do
B200BEV_BENCH_SERIAL=$mode timeout 600 python bench.py --precision f32 --no-cpu-baseline --no-e2e --no-alt --no-configs --steps 15 > gpurun_out/bench_f32_$mode.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/rc.txt
echo "serial=$mode $(python tools/bench_summary.py gpurun_out/bench_f32_$mode.log 2>/dev/null | grep -E '^value' | cut -c1-100)"
done
cat gpurun_out/rc.txt
