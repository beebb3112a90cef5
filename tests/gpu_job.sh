set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt; nproc >> gpurun_out/gpus.txt
tests/run_gpu_groups.sh gpurun_out
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/rc.txt
python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/rc.txt
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "benchref rc=$?" >> gpurun_out/rc.txt
python tests/perf_kernels.py all > gpurun_out/perf_all.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01c.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?" >> gpurun_out/rc.txt
ncu --set full --clock-control none --import-source on -o gpurun_out/prof_r01c_stages -f python tests/prof_stages.py --reps 1 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?" >> gpurun_out/rc.txt
python tests/trace_tc.py gpurun_out/trace_tc.txt > gpurun_out/trace_tc.log 2>&1
ls -la gpurun_out; du -sh gpurun_out
