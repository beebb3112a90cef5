# One gpurun call: parity tests, smoke, bench (both arms), kernel timings, tcgen05 timeline, ncu launch list,
# ncu --set full per stage.  Everything lands in gpurun_out/ (kept under 60 MiB: reports are summarised to CSV
# on the box and dropped if they are too large to travel).
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt; nproc >> gpurun_out/gpus.txt
timeout 500 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/gpu_tests.log 2>&1; echo "pytest -m gpu rc=$?" >> gpurun_out/rc.txt
tail -3 gpurun_out/gpu_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/rc.txt
timeout 600 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/rc.txt
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "benchref rc=$?" >> gpurun_out/rc.txt
timeout 300 python tests/perf_kernels.py all > gpurun_out/perf_all.log 2>&1
timeout 300 python tests/perf_kernels.py all --frames 8 --grid 100 --points 300000 > gpurun_out/perf_stress.log 2>&1
timeout 120 python tests/trace_tc.py gpurun_out/trace_tc.txt > gpurun_out/trace_tc.log 2>&1
if [ "${SKIP_NCU:-0}" != "1" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-alt > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?" >> gpurun_out/rc.txt
ncu --set full --clock-control none -o gpurun_out/prof_stages -f python tests/prof_stages.py --reps 1 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?" >> gpurun_out/rc.txt
python tools/ncu_summary.py gpurun_out/prof_stages.ncu-rep gpurun_out/prof_stages_summary.csv >> gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none -k regex:"linear_rowstream|conv3x3_tc_halo|conv_tc_ws|nchw_to_nhwc" -o gpurun_out/prof_next -f python tests/prof_stages.py --reps 1 --frames 8 --only dense,conv > gpurun_out/ncu_next.log 2>&1; python tools/ncu_summary.py gpurun_out/prof_next.ncu-rep gpurun_out/prof_next_summary.csv >> gpurun_out/ncu_next.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pointnet_mlp_tc -o gpurun_out/prof_tc_cell -f python tests/prof_stages.py --reps 1 --only mlp_tc_cell > gpurun_out/ncu_tc.log 2>&1; echo "ncu tc rc=$?" >> gpurun_out/rc.txt
fi
for f in $(ls -S gpurun_out/*.ncu-rep 2>/dev/null); do
  if [ $(du -sm gpurun_out | cut -f1) -ge 60 ]; then rm -f "$f"; echo "dropped $f (too large to travel)" >> gpurun_out/rc.txt; fi
done
ls -la gpurun_out; du -sh gpurun_out
