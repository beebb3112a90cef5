#!/bin/bash
# GPU job 27: processing time per own layer-5 chunk of an epilogue warp that drains the even chunks (tid 64) / the odd chunks (tid 320)
mkdir -p gpurun_out
for tidv in 64 320; do
  B200BEV_NVCC_EXTRA="-DB200BEV_TC_TRACER_TID=$tidv" python -m bevfusion_multimodal_3d_object_detection_b200.build --force --debug-env > gpurun_out/build_debug.log 2>&1
  timeout 120 python tests/trace_tc.py gpurun_out/trace_tc.txt > gpurun_out/trace_tc.log 2>&1
  cp gpurun_out/trace_tc_cell.txt gpurun_out/trace_tc_cell_tid$tidv.txt
  echo "== tracer tid $tidv"; python tools/tc_set_times.py gpurun_out/trace_tc_cell.txt
done
