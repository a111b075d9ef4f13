mkdir -p gpurun_out
PK_CHOL_CTAS_PER_SM=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/exp_cta1.json 2>> gpurun_out/bench_err.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/exp_cta1.json').read().strip().splitlines()[-1])
print('1 CTA/SM', d['value'], d['roofline']['phase_ms_per_step'])
PY
