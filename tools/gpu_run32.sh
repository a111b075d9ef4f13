mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "split_pipeline or grid_memo or shard_hint or large_batch or golden or tiled_path" > gpurun_out/sanitizer_memcheck_r01_s4.log 2>&1; echo "memcheck rc=$?"; tail -15 gpurun_out/sanitizer_memcheck_r01_s4.log
