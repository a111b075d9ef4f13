timeout 900 python -m pytest tests -m gpu -x -q -k "predict or large_n or full_size_c5 or shard_hint" 2>&1 | tail -4
PK_PROBE_TIME=1 PK_PROBE_REPS=5 timeout 300 python tools/predict_probe.py 2>&1 | tail -2
PK_PROBE_TIME=1 PK_PROBE_REPS=5 timeout 300 python tools/predict_probe.py 4 2>&1 | tail -2
