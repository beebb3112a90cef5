#!/bin/bash
# two GPUs of one box: the driver's launch line for N=2 (and the reference arm under torchrun)
mkdir -p gpurun_out
: > gpurun_out/rc.txt
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 10 --warmup 3 ) > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?" >> gpurun_out/rc.txt
( time timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --impl reference --gpus 2 --steps 2 --warmup 1 ) > gpurun_out/bench_ref_n2.log 2> gpurun_out/bench_ref_n2.err; echo "ref n2 rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/rc.txt
tail -c 600 gpurun_out/bench_n2.err
