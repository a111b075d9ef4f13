timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
PK_CTA_TIMING=1 timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep "pk cta" | tail -1
PK_CTA_TIMING=1 PK_CHOL_CTAS_PER_SM=1 timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep "pk cta" | tail -1
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r01_m.json 2> gpurun_out/bench_err.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r01_m.json').read().strip().splitlines()[-1])
print(d['value'], d['roofline']['phase_ms_per_step'], d['roofline']['frac'])
PY
