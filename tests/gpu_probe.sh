timeout 200 python -m pytest tests -m gpu -q -x -k "random_shapes" -p no:cacheprovider 2>&1 | tail -15
timeout 100 python __graft_entry__.py smoke 2>&1 | tail -2
