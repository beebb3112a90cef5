step() { echo "== $*"; timeout 40 "$@"; rc=$?; echo "rc=$rc"; if [ $rc -eq 124 ]; then echo "HANG: $*"; exit 3; fi; }
step python -m pytest tests -m gpu -q -x -k "tensor_core or cell_canvas or dropin" -p no:cacheprovider 2>&1 | tail -3
[ ${PIPESTATUS[0]} -eq 3 ] && exit 3
for i in 1 2; do timeout 60 python tests/perf_kernels.py mlp 2>&1 | grep "tcgen05" || exit 3; done
timeout 60 python tests/trace_tc.py gpurun_out/trace_tc_pair.txt > gpurun_out/trace_tc.log 2>&1
