mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r01_u.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r01_u.log
tail -12 gpurun_out/pytest_gpu_r01_u.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r01_u.json 2> gpurun_out/bench_err.log; tail -3 gpurun_out/bench_err.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r01_u.json').read().strip().splitlines()[-1])
print(d['value'], d['roofline']['phase_ms_per_step'], d['predictions'])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:predict_reg -c 1 -o gpurun_out/prof_predict_reg_r01_u -f python tools/predict_probe.py > gpurun_out/ncu_u4.log 2>&1; tail -2 gpurun_out/ncu_t4.log
