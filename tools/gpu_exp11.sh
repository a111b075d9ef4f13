mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_models.py -m gpu -x -q -k slsqp 2>&1 | grep -E "assert|Error|passed|failed" | head -20
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:psi_pop -c 1 -o gpurun_out/prof_psi_pop_r01_x -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-predictions > gpurun_out/ncu_x.log 2>&1; tail -2 gpurun_out/ncu_x.log
