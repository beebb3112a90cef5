#!/bin/bash
# GPU job 1 of round 2: GPU test-suite, the new bench (all BASELINE configs), memcheck of the non-tcgen05 kernels.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt; lscpu | head -30 >> gpurun_out/nproc.txt
: > gpurun_out/rc.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/rc.txt
( time timeout 900 python bench.py --steps 10 --warmup 3 ) > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/rc.txt
( time timeout 300 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "ref rc=$?" >> gpurun_out/rc.txt
timeout 300 python tests/sanitize_run.py > gpurun_out/sanitize_plain.log 2>&1 && \
  ( time timeout 900 compute-sanitizer --tool memcheck python tests/sanitize_run.py ) > gpurun_out/sanitize_memcheck.log 2>&1
echo "sanitize rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/rc.txt
tail -5 gpurun_out/gpu_tests.log
tail -c 600 gpurun_out/bench.err
tail -5 gpurun_out/sanitize_memcheck.log
