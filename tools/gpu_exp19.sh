timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/shape_sweep.py 2>&1 | tail -10 | cut -c1-260
