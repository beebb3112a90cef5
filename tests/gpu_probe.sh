timeout 100 python -m pytest tests -m gpu -q -x -k "radar or lidar" -p no:cacheprovider 2>&1 | tail -3
timeout 100 python tests/perf_kernels.py radar 2>&1 | tail -1
timeout 100 python tests/perf_kernels.py mlp 2>&1 | grep f32
