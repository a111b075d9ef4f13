mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r01_w.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r01_w.log
tail -8 gpurun_out/pytest_gpu_r01_w.log
timeout 900 python bench.py > gpurun_out/bench_r01_w.json 2> gpurun_out/bench_err.log; echo "bench rc=$?"; tail -3 gpurun_out/bench_err.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r01_w.json 2>> gpurun_out/bench_err.log; echo "ref rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r01_w.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e'], d['roofline']['frac'], d['roofline']['phase_ms_per_step'], d['predictions'], d['cpu_baseline'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01_w.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_w.log 2>&1; tail -2 gpurun_out/ncu_w.log
