PK_CTA_TIMING=1 timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-predictions 2>&1 >/dev/null | grep "pk cta" | tail -3
PK_CTA_TIMING=1 PK_CHOL_CTAS_PER_SM=1 timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-predictions 2>&1 >/dev/null | grep "pk cta" | tail -3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:chol_cta -c 1 -o gpurun_out/prof_chol_cta_r01_z -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-predictions > gpurun_out/ncu_z1.log 2>&1; tail -2 gpurun_out/ncu_z1.log
