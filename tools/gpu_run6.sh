mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r01_n.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r01_n.log
tail -6 gpurun_out/pytest_gpu_r01_n.log
timeout 600 python bench.py > gpurun_out/bench_r01_n.json 2> gpurun_out/bench_err.log; tail -3 gpurun_out/bench_err.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r01_n.json 2>> gpurun_out/bench_err.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r01_n.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_n1.log 2>&1
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r01_n.json').read().strip().splitlines()[-1])
print(d['value'], d['roofline']['phase_ms_per_step'], d['roofline']['frac'], d['e2e'], d.get('predictions'), d.get('cpu_baseline'))
print(open('gpurun_out/bench_ref_r01_n.json').read()[:300])
PY
