#!/bin/bash
# GPU job 38: tensor-core dense layer at fp32 accuracy (dense_split_tc.cu): parity + timing
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "dense or lidar_init" ) > gpurun_out/gpu_tests_dense.log 2>&1; echo "pytest rc=$?" > gpurun_out/rc.txt
tail -15 gpurun_out/gpu_tests_dense.log
timeout 300 python tests/perf_kernels.py dense > gpurun_out/perf_dense.log 2>&1; echo "perf rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/perf_dense.log | tail -14
cat gpurun_out/rc.txt
