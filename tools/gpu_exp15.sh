timeout 900 python tools/shape_sweep.py 2>&1 | tail -12
