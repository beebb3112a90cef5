#!/bin/bash
# GPU job 3: GPU test-suite (incl. guard bands), then the launch list of one eager step per precision.
mkdir -p gpurun_out
: > gpurun_out/rc.txt
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/rc.txt
timeout 300 python tests/probes/launch_list.py > gpurun_out/launch_plain.log 2>&1 && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_step.csv \
     python tests/probes/launch_list.py > gpurun_out/launch_ncu.log 2>&1
echo "ncu rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/rc.txt
tail -8 gpurun_out/gpu_tests.log
tail -3 gpurun_out/launch_ncu.log
