#!/bin/bash
# GPU job 17: timeline (debug-env build) and ncu --set full of the cell-mode kernel after the no-ack epilogue
mkdir -p gpurun_out
rm -f gpurun_out/prof_tc_cell.ncu-rep
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pointnet_mlp_tc -c 2 -o gpurun_out/prof_tc_cell -f python tests/prof_stages.py --reps 1 --only mlp_tc_cell > gpurun_out/ncu_tc.log 2>&1; echo "ncu tc rc=$?" > gpurun_out/rc.txt
python -m bevfusion_multimodal_3d_object_detection_b200.build --force --debug-env > gpurun_out/build_debug.log 2>&1; echo "build rc=$?" >> gpurun_out/rc.txt
timeout 120 python tests/trace_tc.py gpurun_out/trace_tc.txt > gpurun_out/trace_tc.log 2>&1; echo "trace rc=$?" >> gpurun_out/rc.txt
python tools/tc_timeline.py gpurun_out/trace_tc.txt gpurun_out/trace_tc_cell.txt > gpurun_out/tc_timeline.txt 2>&1
cat gpurun_out/rc.txt; tail -60 gpurun_out/tc_timeline.txt
