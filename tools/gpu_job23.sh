#!/bin/bash
# GPU job 23: split layout kernel staged through shared memory: parity + f32 bench + launch list
mkdir -p gpurun_out
: > gpurun_out/rc.txt
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/rc.txt
grep -E "passed|failed" gpurun_out/gpu_tests.log | tail -2
grep -E "^FAILED|^ERROR" gpurun_out/gpu_tests.log | head
( time timeout 900 python bench.py --precision f32 --no-cpu-baseline --no-e2e --no-configs --no-alt --steps 10 ) > gpurun_out/bench_f32.log 2> gpurun_out/bench_f32.err; echo "bench rc=$?" >> gpurun_out/rc.txt
python tools/bench_summary.py gpurun_out/bench_f32.log 2>/dev/null | head -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_step_e.csv python tests/probes/launch_list.py 32 > gpurun_out/launch_ncu.log 2>&1; echo "ncu list rc=$?" >> gpurun_out/rc.txt
python tools/launch_shares.py gpurun_out/launches_step_e.csv 2>/dev/null | grep -i "nchw\|absmax"
cat gpurun_out/rc.txt
