mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_r01_s4b.log
timeout 900 python bench.py > gpurun_out/bench_r01_s4b.json 2> gpurun_out/bench_err.log; echo "bench rc=$?"; tail -3 gpurun_out/bench_err.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r01_s4b.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['phase_ms_per_step'], d['predictions']['value'], d['predictions']['frac_of_dmma_peak'], d['cpu_baseline']['value'], d['gpu_launches'])
PY
