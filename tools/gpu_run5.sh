mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r01_l.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r01_l.log
tail -4 gpurun_out/pytest_gpu_r01_l.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r01_l.json 2> gpurun_out/bench_err.log; tail -3 gpurun_out/bench_err.log
PK_CHOL_CTAS_PER_SM=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/exp_cta1_l.json 2>> gpurun_out/bench_err.log
python - <<'PY'
import json
for f in ['gpurun_out/bench_r01_l.json','gpurun_out/exp_cta1_l.json']:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d['value'], d['roofline']['phase_ms_per_step'], d['roofline']['frac'])
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:chol_cta -c 1 -o gpurun_out/prof_chol_cta_r01_l -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l2.log 2>&1
