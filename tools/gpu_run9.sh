mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r01_p.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r01_p.log
tail -4 gpurun_out/pytest_gpu_r01_p.log
PK_CTA_TIMING=1 timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-predictions 2>&1 >/dev/null | grep "pk cta" | tail -1
PK_CTA_TIMING=1 PK_CHOL_CTAS_PER_SM=1 timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-predictions 2>&1 >/dev/null | grep "pk cta" | tail -1
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-predictions > gpurun_out/bench_r01_p.json 2> gpurun_out/bench_err.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r01_p.json').read().strip().splitlines()[-1])
print(d['value'], d['roofline']['phase_ms_per_step'], d['roofline']['frac'], d['roofline_psi']['fp64_pipe_frac'])
PY
