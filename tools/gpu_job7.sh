#!/bin/bash
# GPU job 6: fp32-accuracy convolution blocks (split fp16) — parity, bench (f32 / f32_cudnn / bf16 timed).
mkdir -p gpurun_out
: > gpurun_out/rc.txt
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/rc.txt
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/rc.txt
grep -E "passed|failed" gpurun_out/gpu_tests.log | tail -3
grep -E "^FAILED|^ERROR" gpurun_out/gpu_tests.log | head -20
tail -c 400 gpurun_out/bench.err
