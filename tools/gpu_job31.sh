#!/bin/bash
# GPU job 31: stress workload, parallel branches on / off on one box
mkdir -p gpurun_out
for mode in 0 1 0 1; do
  B200BEV_BENCH_SERIAL=$mode timeout 600 python bench.py --workload stress --no-cpu-baseline --no-e2e --no-alt --no-configs --steps 20 > gpurun_out/bench_stress_$mode.log 2> gpurun_out/bench_stress.err
  echo "serial=$mode $(python tools/bench_summary.py gpurun_out/bench_stress_$mode.log 2>/dev/null | grep -E '^value' | cut -c1-80)"
  python tools/bench_summary.py gpurun_out/bench_stress_$mode.log 2>/dev/null | grep -E '^eager' | cut -c1-220
done
