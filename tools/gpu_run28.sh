mkdir -p gpurun_out
timeout 600 python tools/variant_sweep.py 2>&1 | tee gpurun_out/variant_sweep_r01.jsonl | cut -c1-160
