timeout 60 python tests/_bs_warm.py 2>&1 | tail -12
