#!/bin/bash
# GPU job 37: ncu --set full of the convolution / layout / resize kernels of the bf16 path on the final code (32 frames; CTA-pair 3x3, TMA-fed 1x1)
mkdir -p gpurun_out
rm -f gpurun_out/prof_conv.ncu-rep
timeout 900 ncu --set full --clock-control none -k regex:"conv3x3_tc_halo|conv_tc_ws|conv1x1_tma|nchw_to_nhwc|bilinear_resize_nhwc|camera_mean_nhwc" -o gpurun_out/prof_conv -f python tests/prof_stages.py --reps 1 --frames 32 --only conv > gpurun_out/ncu_conv.log 2>&1; echo "ncu rc=$?" > gpurun_out/rc.txt
python tools/ncu_summary.py gpurun_out/prof_conv.ncu-rep gpurun_out/prof_conv_summary.csv >> gpurun_out/ncu_conv.log 2>&1
rm -f gpurun_out/prof_conv.ncu-rep
cat gpurun_out/rc.txt; tail -3 gpurun_out/ncu_conv.log; cut -d, -f1,5,9,6,7 gpurun_out/prof_conv_summary.csv | head -30
