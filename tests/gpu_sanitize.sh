# compute-sanitizer memcheck over the GPU parity tests of the non-tensor-core kernels (one tool per gpurun call).
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/memcheck.log \
  python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider \
  -k "prepare or bin_sort or radar or camera_mean or resize or nms or topk or hand or decode_small or lidar_global_ragged" \
  > gpurun_out/memcheck_pytest.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/rc.txt
tail -3 gpurun_out/memcheck_pytest.log; tail -5 gpurun_out/memcheck.log; cat gpurun_out/rc.txt
