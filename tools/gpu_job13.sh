#!/bin/bash
# GPU job 13: cell-mode epilogue of the bf16 MLP as a dual-block run walk: parity tests + timings (base and stress shape)
mkdir -p gpurun_out
: > gpurun_out/rc.txt
( time timeout 900 python -m pytest tests -m gpu -q -x ) > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/rc.txt
timeout 300 python tests/perf_kernels.py mlp > gpurun_out/perf_mlp.log 2>&1; echo "perf rc=$?" >> gpurun_out/rc.txt
timeout 300 python tests/perf_kernels.py mlp --frames 8 --grid 100 --points 300000 > gpurun_out/perf_mlp_stress.log 2>&1
cat gpurun_out/rc.txt
grep -E "passed|failed" gpurun_out/gpu_tests.log | tail -3
grep -E "^FAILED|^ERROR|Error|assert" gpurun_out/gpu_tests.log | head -20
grep bf16 gpurun_out/perf_mlp.log gpurun_out/perf_mlp_stress.log
