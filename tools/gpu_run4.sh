set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r01_k.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r01_k.log
tail -5 gpurun_out/pytest_gpu_r01_k.log
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r01_k.json 2> gpurun_out/bench_err.log; tail -3 gpurun_out/bench_err.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:psi_pop -c 1 -o gpurun_out/prof_psi_pop_r01_k -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_k3.log 2>&1
cut -c1-300 gpurun_out/bench_r01_k.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:chol_cta -c 1 -o gpurun_out/prof_chol_cta_r01_k -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_k2.log 2>&1
