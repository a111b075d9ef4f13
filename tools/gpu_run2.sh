set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r01_i.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r01_i.log
tail -15 gpurun_out/pytest_gpu_r01_i.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r01_i_cta.json 2> gpurun_out/bench_err.log; tail -3 gpurun_out/bench_err.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r01_i.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_i1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:chol_cta -c 1 -o gpurun_out/prof_chol_cta_r01_i -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_i2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:psi_pop -c 1 -o gpurun_out/prof_psi_pop_r01_i -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_i3.log 2>&1
cat gpurun_out/bench_r01_i_cta.json | cut -c1-400
