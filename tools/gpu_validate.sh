# Full GPU validation: pytest -m gpu, smoke(), default bench, shape sweep (run under gpurun)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_r01_s4.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_r01_s4.json 2> gpurun_out/bench_err.log; echo "bench rc=$?"; tail -3 gpurun_out/bench_err.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r01_s4.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['predictions']['value'], d['cpu_baseline']['value'], d['gpu_launches'])
PY
timeout 300 python tools/shape_sweep.py 2>&1 | tail -10 | cut -c1-250 | tee gpurun_out/shape_sweep_r01_s4.log
