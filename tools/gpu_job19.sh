#!/bin/bash
mkdir -p gpurun_out
for v in 0 1 2; do
  echo "== variant $v (debug-env build)"
  B200BEV_NVCC_EXTRA="-DB200BEV_SCOUT_CHECK -DB200BEV_VARIANT=$v" python -m bevfusion_multimodal_3d_object_detection_b200.build --force --debug-env > gpurun_out/build_x.log 2>&1
  timeout 300 python tests/perf_kernels.py mlp 2>&1 | grep -v "f32" | tail -6 | cut -c1-200
done
