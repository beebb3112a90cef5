timeout 100 python -m pytest tests -m gpu -q -x -k "camera" -p no:cacheprovider 2>&1 | tail -3
timeout 100 python tests/perf_kernels.py camera 2>&1 | tail -3
timeout 100 python tests/perf_kernels.py camera --grid 100 2>&1 | tail -3
