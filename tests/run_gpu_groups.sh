#!/bin/bash
# Runs the GPU parity tests group by group with hard timeouts, so one hung kernel cannot eat the box.
# usage: tests/run_gpu_groups.sh [outdir]
out=${1:-gpurun_out}
mkdir -p "$out"
for grp in library prepare bin_sort lidar tensor_core canvas radar camera nms_and topk hand decode module chain; do
  timeout -k 5 ${GROUP_TIMEOUT:-120} python -X faulthandler -m pytest tests -m gpu -q -x -k "$grp" -p no:cacheprovider \
      --timeout 100 > "$out/gpu_$grp.log" 2>&1
  echo "$grp rc=$?" >> "$out/gpu_groups.txt"
done
cat "$out/gpu_groups.txt"
