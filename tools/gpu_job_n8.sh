#!/bin/bash
# eight GPUs of one box: the driver's launch line for N=8
mkdir -p gpurun_out
: > gpurun_out/rc.txt
nvidia-smi topo -m > gpurun_out/topo8.txt 2>&1
nproc >> gpurun_out/topo8.txt; numactl -H >> gpurun_out/topo8.txt 2>&1; lscpu | grep -i -E "numa|socket|model name|^CPU\(s\)" >> gpurun_out/topo8.txt
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 \
    bench.py --gpus 8 --steps 10 --warmup 3 ) > gpurun_out/bench_n8.log 2> gpurun_out/bench_n8.err; echo "bench n8 rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/rc.txt
tail -c 600 gpurun_out/bench_n8.err
