#!/bin/bash
# GPU job 42: evidence set on the committed code (tensor-core dense layer, CTA-pair convolution and fp32-accuracy MLP): GPU suite, smoke, default bench (both arms),
# kernel timings (bench / stress shapes), ncu launch list of the step
mkdir -p gpurun_out
: > gpurun_out/rc.txt
( time timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider ) > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/rc.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/rc.txt
( time timeout 900 python bench.py ) > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/rc.txt
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "benchref rc=$?" >> gpurun_out/rc.txt
timeout 300 python tests/perf_kernels.py all > gpurun_out/perf_all.log 2>&1
timeout 300 python tests/perf_kernels.py all --frames 8 --grid 100 --points 300000 > gpurun_out/perf_stress.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_step_m.csv python tests/probes/launch_list.py 32 > gpurun_out/launch_ncu.log 2>&1; echo "ncu list rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/rc.txt
grep -E "passed|failed" gpurun_out/gpu_tests.log | tail -2
tail -2 gpurun_out/smoke.log
python tools/bench_summary.py gpurun_out/bench.log 2>/dev/null | head -40
cat gpurun_out/bench_ref.log | cut -c1-600
