#!/bin/bash
# GPU job 4: per-warp transposing cell epilogue — parity, bench, ncu of the cell-mode kernel.
mkdir -p gpurun_out
: > gpurun_out/rc.txt
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/rc.txt
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/rc.txt
timeout 300 python tests/probes/launch_list.py 32 > gpurun_out/launch_plain.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:pointnet_mlp_tc_kernel -c 2 -o gpurun_out/prof_tc_cell \
     python tests/probes/launch_list.py 32 > gpurun_out/ncu_tc.log 2>&1
echo "ncu rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/rc.txt
tail -6 gpurun_out/gpu_tests.log
tail -c 400 gpurun_out/bench.err
