mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r01_r.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r01_r.log
tail -3 gpurun_out/pytest_gpu_r01_r.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-predictions > gpurun_out/bench_r01_r.json 2> gpurun_out/bench_err.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r01_r.json').read().strip().splitlines()[-1])
print(d['value'], d['roofline']['phase_ms_per_step'], d['roofline']['frac'], d['roofline_psi']['fp64_pipe_frac'])
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:psi_pop -c 1 -o gpurun_out/prof_psi_pop_r01_r -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-predictions > gpurun_out/ncu_r3.log 2>&1
