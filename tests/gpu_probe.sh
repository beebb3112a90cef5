timeout 100 python -m pytest tests -m gpu -q -x -k "prepare" -p no:cacheprovider 2>&1 | tail -5
timeout 60 python tests/perf_kernels.py prepare 2>&1 | tail -1
