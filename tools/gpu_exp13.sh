PK_CTA_TIMING=1 timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-predictions 2>&1 >/dev/null | grep "pk cta" | tail -3
PK_CTA_TIMING=1 PK_CHOL_CTAS_PER_SM=1 timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-predictions 2>&1 >/dev/null | grep "pk cta" | tail -3
