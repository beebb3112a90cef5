#!/bin/bash
# GPU job 45: ncu --set full of the tensor-core dense layer (and the FFMA streaming kernel beside it), batch 32 and 1
mkdir -p gpurun_out
rm -f gpurun_out/prof_dense.ncu-rep
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"dense_split_tc|linear_stream|linear_rowstream" -c 8 -o gpurun_out/prof_dense -f python tests/prof_dense.py 32 1 > gpurun_out/ncu_dense.log 2>&1; echo "ncu rc=$?" > gpurun_out/rc.txt
python tools/ncu_summary.py gpurun_out/prof_dense.ncu-rep gpurun_out/prof_dense_summary.csv >> gpurun_out/ncu_dense.log 2>&1
ncu -i gpurun_out/prof_dense.ncu-rep --page details --csv 2>/dev/null | grep -E "dense_split_tc" | grep -E "Stall|Issued|Eligible|No Eligible|Active Warps|DRAM Throughput|Memory Throughput|L2 Hit" | cut -d, -f5,12-15 | head -40 > gpurun_out/prof_dense_details.txt
ls -la gpurun_out/prof_dense.ncu-rep; if [ $(stat -c %s gpurun_out/prof_dense.ncu-rep) -gt 30000000 ]; then rm -f gpurun_out/prof_dense.ncu-rep; fi
cat gpurun_out/rc.txt; tail -3 gpurun_out/ncu_dense.log; cat gpurun_out/prof_dense_summary.csv | cut -c1-400
