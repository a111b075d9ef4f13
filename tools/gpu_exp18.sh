PK_CTA_TIMING=1 timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-predictions 2>&1 >/dev/null | grep "pk cta" | tail -3
PK_CTA_TIMING=1 PK_CHOL_CTAS_PER_SM=1 timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-predictions 2>&1 >/dev/null | grep "pk cta" | tail -3
PK_CHOL_CTAS_PER_SM=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-predictions 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('1cta/sm', d['value'], d['roofline']['phase_ms_per_step'])"
