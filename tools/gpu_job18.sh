#!/bin/bash
# GPU job 18: scout warp + l5_done: full parity suite, timings, timeline (debug-env build)
mkdir -p gpurun_out
: > gpurun_out/rc.txt
( time timeout 900 python -m pytest tests -m gpu -q -x  ) > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/rc.txt
grep -E "passed|failed" gpurun_out/gpu_tests.log | tail -3
grep -E "^FAILED|^ERROR|Error|assert" gpurun_out/gpu_tests.log | head -20
: > gpurun_out/perf_mlp.log
for i in 1 2; do timeout 300 python tests/perf_kernels.py mlp 2>&1 | grep bf16 >> gpurun_out/perf_mlp.log; done
timeout 300 python tests/perf_kernels.py mlp --frames 8 --grid 100 --points 300000 2>&1 | grep bf16 >> gpurun_out/perf_mlp.log
python -m bevfusion_multimodal_3d_object_detection_b200.build --force --debug-env > gpurun_out/build_debug.log 2>&1; echo "build rc=$?" >> gpurun_out/rc.txt
timeout 120 python tests/trace_tc.py gpurun_out/trace_tc.txt > gpurun_out/trace_tc.log 2>&1; echo "trace rc=$?" >> gpurun_out/rc.txt
python tools/tc_timeline.py gpurun_out/trace_tc_cell.txt > gpurun_out/tc_timeline.txt 2>&1
cat gpurun_out/perf_mlp.log; cat gpurun_out/rc.txt; cat gpurun_out/tc_timeline.txt
