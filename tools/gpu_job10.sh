#!/bin/bash
# GPU job 10: launch list of the fp32 step and of the fp32-accuracy MLP (global / cell)
mkdir -p gpurun_out
: > gpurun_out/rc.txt
timeout 300 python tests/probes/launch_list.py 32 > gpurun_out/launch_plain.log 2>&1 && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_step_b.csv \
     python tests/probes/launch_list.py 32 > gpurun_out/launch_ncu.log 2>&1
echo "ncu rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/rc.txt; tail -3 gpurun_out/launch_ncu.log
