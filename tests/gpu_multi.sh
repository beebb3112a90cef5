mkdir -p gpurun_out
N=${1:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err; echo "bench n$N rc=$?" >> gpurun_out/rc.txt
tail -1 gpurun_out/bench_n$N.log | cut -c1-400; tail -5 gpurun_out/bench_n$N.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_n$N.log 2> gpurun_out/bench_ref_n$N.err; echo "benchref n$N rc=$?" >> gpurun_out/rc.txt
tail -1 gpurun_out/bench_ref_n$N.log | cut -c1-300
cat gpurun_out/rc.txt
