#!/bin/bash
# GPU job 11: full GPU test-suite + smoke on the current tree (cell ids of the split kernel's epilogue prefetched), perf of every kernel
mkdir -p gpurun_out
: > gpurun_out/rc.txt
( time timeout 900 python -m pytest tests -m gpu -q -x ) > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/rc.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/rc.txt
timeout 300 python tests/perf_kernels.py all > gpurun_out/perf_all.log 2>&1; echo "perf rc=$?" >> gpurun_out/rc.txt
timeout 300 python tests/probes/split_probe.py > gpurun_out/split_probe.log 2>&1; echo "split probe rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/rc.txt
grep -E "passed|failed" gpurun_out/gpu_tests.log | tail -3
tail -30 gpurun_out/perf_all.log
tail -12 gpurun_out/split_probe.log
