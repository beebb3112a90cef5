#!/bin/bash
# GPU job 49: the head's five outputs by one fused copy launch: head / chain tests, short bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_dropin.py tests/test_gpu_bench_step.py -m gpu -q -x -p no:cacheprovider -k "head or chain or decode or graphed or step" 2>&1 | tail -3
timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-configs --no-alt > gpurun_out/bench_short.log 2> gpurun_out/bench_short.err; echo "bench rc=$?"
python tools/bench_summary.py gpurun_out/bench_short.log 2>/dev/null | head -3
