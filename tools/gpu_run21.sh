mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "predict or large_n" 2>&1 | tail -8
