#!/bin/bash
# GPU job 16: cell-mode epilogue without acknowledgements (l5_done barrier, global maxima by atomics): parity, timings;
# then the MMA issuer's chunk loop unrolled in cell mode (experiment build)
mkdir -p gpurun_out
: > gpurun_out/rc.txt
( time timeout 900 python -m pytest tests -m gpu -q -x ) > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/rc.txt
grep -E "passed|failed" gpurun_out/gpu_tests.log | tail -3
grep -E "^FAILED|^ERROR|Error|assert" gpurun_out/gpu_tests.log | head -20
: > gpurun_out/perf_mlp.log
for i in 1 2; do timeout 300 python tests/perf_kernels.py mlp 2>&1 | grep bf16 >> gpurun_out/perf_mlp.log; done
timeout 300 python tests/perf_kernels.py mlp --frames 8 --grid 100 --points 300000 2>&1 | grep bf16 >> gpurun_out/perf_mlp.log
echo "== MMA chunk loop unrolled" >> gpurun_out/perf_mlp.log
B200BEV_NVCC_EXTRA="-DB200BEV_CELL_MMA_UNROLL=1" python -m bevfusion_multimodal_3d_object_detection_b200.build --force > gpurun_out/build_x.log 2>&1; echo "build rc=$?" >> gpurun_out/rc.txt
for i in 1 2; do timeout 300 python tests/perf_kernels.py mlp 2>&1 | grep bf16 >> gpurun_out/perf_mlp.log; done
timeout 300 python tests/perf_kernels.py mlp --frames 8 --grid 100 --points 300000 2>&1 | grep bf16 >> gpurun_out/perf_mlp.log
cat gpurun_out/perf_mlp.log; cat gpurun_out/rc.txt
