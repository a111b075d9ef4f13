# Final round-2 validation on one GPU: pytest -m gpu, smoke(), bench (both arms), launch list, ncu --set full captures
mkdir -p gpurun_out
T=${1:-u}
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_gpu_r02_$T.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_r02_$T.json 2> gpurun_out/bench_err.log; echo "bench rc=$?"; tail -3 gpurun_out/bench_err.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r02_$T.json 2> gpurun_out/bench_ref_err.log; echo "ref rc=$?"; cut -c1-400 gpurun_out/bench_ref_r02_$T.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02_$T.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$T.log 2>&1; tail -1 gpurun_out/ncu_$T.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:chol_i8 -c 1 -o gpurun_out/prof_chol_i8_r02_$T -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-predictions --no-extras > gpurun_out/ncu_${T}1.log 2>&1; tail -1 gpurun_out/ncu_${T}1.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:psi_pop -c 1 -o gpurun_out/prof_psi_pop_r02_$T -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-predictions --no-extras > gpurun_out/ncu_${T}2.log 2>&1; tail -1 gpurun_out/ncu_${T}2.log | cut -c1-200
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r02_$T.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline'], d['predictions']['value'], d['cpu_baseline'], d['gpu_launches'])
for k in ('pop300','c4','e2e_api','train_wall_s','parity_sample','roofline_psi','clocks'):
    print(k, json.dumps(d.get(k)))
PY
