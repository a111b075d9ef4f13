timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_models.py -m gpu -x -q 2>&1 | tail -4
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/err.log | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['predictions']['value'], d['predictions']['ms'])" || tail -5 gpurun_out/err.log
