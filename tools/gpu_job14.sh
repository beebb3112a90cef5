#!/bin/bash
# GPU job 14: what the canvas stores of the cell-mode epilogue cost: debug-env build, atomics replaced by plain stores (8),
# no stores (16); with and without the canvas memset in front (perf_kernels times the C-ABI call, memset included)
mkdir -p gpurun_out
python -m bevfusion_multimodal_3d_object_detection_b200.build --force --debug-env > gpurun_out/build_debug.log 2>&1; echo "build rc=$?" > gpurun_out/rc.txt
: > gpurun_out/cell_experiments2.log
for d in 0 8 16 2; do
  echo "== B200BEV_TC_DEBUG=$d" >> gpurun_out/cell_experiments2.log
  B200BEV_TC_DEBUG=$d timeout 300 python tests/perf_kernels.py mlp 2>&1 | grep "bf16 tcgen05 canvas" >> gpurun_out/cell_experiments2.log
done
for d in 0 8 16; do
  echo "== stress B200BEV_TC_DEBUG=$d" >> gpurun_out/cell_experiments2.log
  B200BEV_TC_DEBUG=$d timeout 300 python tests/perf_kernels.py mlp --frames 8 --grid 100 --points 300000 2>&1 | grep "bf16 tcgen05 canvas" >> gpurun_out/cell_experiments2.log
done
cat gpurun_out/cell_experiments2.log; cat gpurun_out/rc.txt
