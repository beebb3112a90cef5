#!/bin/bash
# GPU job 36: fp32-accuracy MLP, run walk of the final layer with four-slot groups and value-independent stores: parity + timing
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?" > gpurun_out/rc.txt
grep -E "passed|failed" gpurun_out/gpu_tests.log | tail -2; grep -E "^FAILED|^ERROR" gpurun_out/gpu_tests.log | head
timeout 300 python tests/probes/split_probe.py 2>&1 | tail -5
cat gpurun_out/rc.txt
