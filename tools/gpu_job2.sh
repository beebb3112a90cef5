#!/bin/bash
# GPU job 2: the split (fp32-accuracy tensor-core) MLP — layer-by-layer probe, then the GPU test-suite.
mkdir -p gpurun_out
: > gpurun_out/rc.txt
( time timeout 300 python tests/probes/split_probe.py ) > gpurun_out/split_probe.log 2>&1; echo "probe rc=$?" >> gpurun_out/rc.txt
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/rc.txt
tail -40 gpurun_out/split_probe.log
tail -15 gpurun_out/gpu_tests.log
