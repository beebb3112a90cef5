#!/bin/bash
# GPU job 44: convolution timings with the channels-last bf16 output only (what the step runs)
mkdir -p gpurun_out
timeout 300 python tests/perf_kernels.py conv > gpurun_out/perf_conv.log 2>&1; echo "perf rc=$?" > gpurun_out/rc.txt
grep -E "channels-last|total" gpurun_out/perf_conv.log; cat gpurun_out/rc.txt
