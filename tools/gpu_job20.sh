#!/bin/bash
# GPU job 20: full GPU suite + smoke on the committed kernels, default bench line, per-kernel timings,
# ncu launch list of the step and ncu --set full of the cell-mode MLP kernel
mkdir -p gpurun_out
: > gpurun_out/rc.txt
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/rc.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/rc.txt
( time timeout 900 python bench.py ) > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/rc.txt
timeout 300 python tests/perf_kernels.py all > gpurun_out/perf_all.log 2>&1
timeout 300 python tests/perf_kernels.py all --frames 8 --grid 100 --points 300000 > gpurun_out/perf_stress.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_step_c.csv python tests/probes/launch_list.py 32 > gpurun_out/launch_ncu.log 2>&1; echo "ncu list rc=$?" >> gpurun_out/rc.txt
rm -f gpurun_out/prof_tc_cell.ncu-rep
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pointnet_mlp_tc -c 2 -o gpurun_out/prof_tc_cell -f python tests/prof_stages.py --reps 1 --only mlp_tc_cell > gpurun_out/ncu_tc.log 2>&1; echo "ncu tc rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/rc.txt
grep -E "passed|failed" gpurun_out/gpu_tests.log | tail -2
tail -c 600 gpurun_out/bench.err
grep bf16 gpurun_out/perf_all.log gpurun_out/perf_stress.log
