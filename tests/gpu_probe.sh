timeout 300 python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -4
timeout 100 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 200 python bench.py --no-cpu-baseline --steps 20 > gpurun_out/bench_pair.log 2>&1; tail -1 gpurun_out/bench_pair.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d['e2e']['value'])
for k,v in d['kernels'].items(): print(k, v['ms'], v['frac'])
"
for g in 100; do timeout 100 python tests/perf_kernels.py mlp --frames 8 --grid 100 --points 300000 2>&1 | grep tcgen05; done
