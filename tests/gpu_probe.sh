timeout 100 python -m pytest tests -m gpu -q -x -k "bin_sort or cell_canvas" -p no:cacheprovider 2>&1 | tail -3
timeout 100 python tests/perf_kernels.py binsort 2>&1 | tail -1
timeout 100 python tests/perf_kernels.py binsort --frames 8 --grid 100 --points 300000 2>&1 | tail -1
timeout 100 python tests/perf_kernels.py binsort --frames 32 --grid 100 --points 300000 2>&1 | tail -1
