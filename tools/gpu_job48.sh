#!/bin/bash
# GPU job 48: CTA-pair form of the fp32-accuracy MLP GEMM (split_gemm_kernel<FINAL, 2>): parity under a short timeout, then timings
mkdir -p gpurun_out
( time timeout 240 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py -m gpu -q -x -p no:cacheprovider -k "pointnet or canvas or lidar_encoder or split or detector_chain or encode_fuse" ) > gpurun_out/gpu_tests_split.log 2>&1; echo "pytest rc=$?" > gpurun_out/rc.txt
tail -8 gpurun_out/gpu_tests_split.log
if grep -q "pytest rc=0" gpurun_out/rc.txt; then
  timeout 200 python tests/probes/split_probe.py > gpurun_out/split_probe.log 2>&1; echo "probe rc=$?" >> gpurun_out/rc.txt
  tail -6 gpurun_out/split_probe.log
  timeout 300 python bench.py --precision f32 --no-cpu-baseline --no-e2e --no-configs --no-alt --steps 10 > gpurun_out/bench_f32.log 2> gpurun_out/bench_f32.err; echo "bench rc=$?" >> gpurun_out/rc.txt
  python tools/bench_summary.py gpurun_out/bench_f32.log 2>/dev/null | head -3
fi
cat gpurun_out/rc.txt
