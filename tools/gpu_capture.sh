# Bench + launch list + ncu --set full captures of the four hot kernels (run under gpurun; summaries -> profiles/ with tools/summarize_ncu.py)
mkdir -p gpurun_out
T=s4
timeout 900 python bench.py > gpurun_out/bench_r01_$T.json 2> gpurun_out/bench_err.log; echo "bench rc=$?"; tail -3 gpurun_out/bench_err.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r01_$T.json 2>> gpurun_out/bench_err.log; echo "ref rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r01_s4.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['phase_ms_per_step'], d['predictions']['value'], d['predictions']['frac_of_dmma_peak'], d['cpu_baseline'], d['gpu_launches'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01_$T.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$T.log 2>&1; tail -1 gpurun_out/ncu_$T.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:chol_cta -c 1 -o gpurun_out/prof_chol_cta_r01_$T -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-predictions > gpurun_out/ncu_${T}1.log 2>&1; tail -1 gpurun_out/ncu_${T}1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:psi_pop -c 1 -o gpurun_out/prof_psi_pop_r01_$T -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-predictions > gpurun_out/ncu_${T}2.log 2>&1; tail -1 gpurun_out/ncu_${T}2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trinorm -c 1 -s 2 -o gpurun_out/prof_trinorm_r01_$T -f python tools/predict_probe.py > gpurun_out/ncu_${T}3.log 2>&1; tail -1 gpurun_out/ncu_${T}3.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:psi_rows_grid -c 1 -s 20 -o gpurun_out/prof_psi_rows_grid_r01_$T -f env PK_PROBE_TIME=1 PK_PROBE_REPS=1 python tools/predict_probe.py > gpurun_out/ncu_${T}4.log 2>&1; tail -1 gpurun_out/ncu_${T}4.log
