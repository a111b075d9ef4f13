timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python tools/shape_sweep.py 2>&1 | tail -8 | cut -c1-250
