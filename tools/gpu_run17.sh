mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_r01_z.json 2> gpurun_out/bench_err.log; echo "bench rc=$?"; tail -3 gpurun_out/bench_err.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r01_z.json 2>> gpurun_out/bench_err.log; echo "ref rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r01_z.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e'], d['roofline']['frac'], d['roofline']['phase_ms_per_step'], d['predictions']['value'], d['cpu_baseline'], d['gpu_launches'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01_z.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_z.log 2>&1; tail -2 gpurun_out/ncu_z.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:chol_cta -c 1 -o gpurun_out/prof_chol_cta_r01_z -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-predictions > gpurun_out/ncu_z1.log 2>&1; tail -2 gpurun_out/ncu_z1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:psi_pop -c 1 -o gpurun_out/prof_psi_pop_r01_z -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-predictions > gpurun_out/ncu_z2.log 2>&1; tail -2 gpurun_out/ncu_z2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:predict_reg -c 1 -o gpurun_out/prof_predict_reg_r01_z -f python tools/predict_probe.py > gpurun_out/ncu_z3.log 2>&1; tail -2 gpurun_out/ncu_z3.log
