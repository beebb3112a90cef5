#!/bin/bash
# GPU job 46: CTA-pair (cta_group::2) form of the 3x3 convolution kernel: parity under a short timeout, then timings
mkdir -p gpurun_out
( time timeout 150 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "conv_bn_relu_tcgen05 or conv_split or conv_bn_relu_split" ) > gpurun_out/gpu_tests_conv.log 2>&1; echo "pytest rc=$?" > gpurun_out/rc.txt
tail -12 gpurun_out/gpu_tests_conv.log
if grep -q "pytest rc=0" gpurun_out/rc.txt; then
  timeout 200 python tests/perf_kernels.py conv > gpurun_out/perf_conv.log 2>&1; echo "perf rc=$?" >> gpurun_out/rc.txt
  grep -E "channels-last|total" gpurun_out/perf_conv.log
fi
cat gpurun_out/rc.txt
