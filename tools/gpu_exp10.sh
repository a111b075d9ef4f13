mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-predictions 2>gpurun_out/err.log | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['phase_ms_per_step'])" || tail -5 gpurun_out/err.log
