#!/bin/bash
# GPU job 15: ncu --set full + source counters of the cell-mode kernel (dual-block walk)
mkdir -p gpurun_out
rm -f gpurun_out/prof_tc_cell.ncu-rep
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pointnet_mlp_tc -c 2 -o gpurun_out/prof_tc_cell -f python tests/prof_stages.py --reps 1 --only mlp_tc_cell > gpurun_out/ncu_tc.log 2>&1; echo "ncu tc rc=$?" > gpurun_out/rc.txt
tail -3 gpurun_out/ncu_tc.log; cat gpurun_out/rc.txt; ls -la gpurun_out/*.ncu-rep
