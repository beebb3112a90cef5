mkdir -p gpurun_out
timeout 120 tests/cuda/_build/umma_rate > gpurun_out/umma_rate.log 2>&1; echo "umma_rate rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/umma_rate.log
