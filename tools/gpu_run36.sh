timeout 300 python tools/fit_probe.py 2>&1 | tail -8
timeout 900 python -m pytest tests -m gpu -x -q -k "fit or predict or models or golden or attributes or large_n" 2>&1 | tail -4
