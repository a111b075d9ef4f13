timeout 600 python -m pytest tests -m gpu -x -q -k "cholesky_variants" 2>&1 | tail -6
PK_SWEEP_SMALL=1 timeout 600 python tools/variant_sweep.py 2>&1 | tee gpurun_out/variant_sweep_small_r01.jsonl | cut -c1-200
