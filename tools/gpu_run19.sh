mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu_r01_s4a.log; cat gpurun_out/pytest_gpu_r01_s4a.log
timeout 300 python tools/shape_sweep.py 2>&1 | tail -12 | cut -c1-300 | tee gpurun_out/shape_sweep_s4a.log
