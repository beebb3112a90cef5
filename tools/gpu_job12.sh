#!/bin/bash
# GPU job 12: (a) tensor-memory fragment-layout probe; (b) what bounds the cell-mode epilogue of the bf16 MLP:
# debug-env build, the kernel timed with the transposing stores / loads skipped (8), the run walk skipped (16), both (24),
# the whole layer-5 epilogue skipped (2)
mkdir -p gpurun_out
: > gpurun_out/rc.txt
timeout 120 tests/cuda/_build/tmem_shapes_probe > gpurun_out/tmem_shapes_probe.log 2>&1; echo "probe rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/tmem_shapes_probe.log | cut -c1-400
python -m bevfusion_multimodal_3d_object_detection_b200.build --force --debug-env > gpurun_out/build_debug.log 2>&1; echo "build rc=$?" >> gpurun_out/rc.txt
: > gpurun_out/cell_experiments.log
for d in 0 8 16 24 2; do
  echo "== B200BEV_TC_DEBUG=$d" >> gpurun_out/cell_experiments.log
  B200BEV_TC_DEBUG=$d timeout 300 python tests/perf_kernels.py mlp 2>&1 | grep bf16 >> gpurun_out/cell_experiments.log
done
for d in 0 8 24; do
  echo "== stress B200BEV_TC_DEBUG=$d" >> gpurun_out/cell_experiments.log
  B200BEV_TC_DEBUG=$d timeout 300 python tests/perf_kernels.py mlp --frames 8 --grid 100 --points 300000 2>&1 | grep bf16 >> gpurun_out/cell_experiments.log
done
cat gpurun_out/cell_experiments.log; cat gpurun_out/rc.txt
