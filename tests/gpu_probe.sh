timeout 40 tests/cuda/_build/umma2_probe; echo "rc=$?"
