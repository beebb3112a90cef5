#!/bin/bash
# GPU job 26: default bench (e2e pipeline included) with the forked branches
mkdir -p gpurun_out
( time timeout 900 python bench.py --no-cpu-baseline ) > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" > gpurun_out/rc.txt
python tools/bench_summary.py gpurun_out/bench.log 2>/dev/null | head -7 | cut -c1-400
cat gpurun_out/rc.txt; tail -c 300 gpurun_out/bench.err
