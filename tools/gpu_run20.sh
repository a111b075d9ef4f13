mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "predict or large_n or golden or models" 2>&1 | tail -5
PK_PROBE_TIME=1 timeout 300 python tools/predict_probe.py 2>&1 | tail -3
