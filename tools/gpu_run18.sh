mkdir -p gpurun_out
nvidia-smi -L
python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/dist_check.py > gpurun_out/dist_check_2gpu_r01_y.log 2>&1; echo "dist rc=$?"; tail -5 gpurun_out/dist_check_2gpu_r01_y.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_r01_y_2gpu.json 2> gpurun_out/bench2_err.log; echo "bench2 rc=$?"; tail -c 600 gpurun_out/bench_r01_y_2gpu.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 2>/dev/null | cut -c1-200
