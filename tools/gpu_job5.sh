#!/bin/bash
# GPU job 5: radar-branch shortcut + border_expand — parity and bench.
mkdir -p gpurun_out
: > gpurun_out/rc.txt
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/rc.txt
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/rc.txt
python -c "from __graft_entry__ import smoke; smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/rc.txt
tail -6 gpurun_out/gpu_tests.log
tail -c 400 gpurun_out/bench.err; tail -3 gpurun_out/smoke.log
