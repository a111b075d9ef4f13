set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r01_h.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r01_h.log
tail -15 gpurun_out/pytest_gpu_r01_h.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r01_h_cta.json 2> gpurun_out/bench_err.log; tail -3 gpurun_out/bench_err.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --chol-variant 1 > gpurun_out/bench_r01_h_steps.json 2>> gpurun_out/bench_err.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r01_h.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_h1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:chol_cta -c 1 -o gpurun_out/prof_chol_cta_r01_h -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_h2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:psi_pop -c 1 -o gpurun_out/prof_psi_pop_r01_h -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_h3.log 2>&1
cat gpurun_out/bench_r01_h_cta.json | cut -c1-400
