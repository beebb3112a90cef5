timeout 200 python -m pytest tests/test_gpu_dropin.py -m gpu -q -p no:cacheprovider 2>&1 | tail -25
