mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/dist_check_r01.log 2>&1; echo "rc=$?" >> gpurun_out/dist_check_r01.log; tail -5 gpurun_out/dist_check_r01.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_r01_n_2gpu.json 2> gpurun_out/bench2_err.log; tail -3 gpurun_out/bench2_err.log; cut -c1-200 gpurun_out/bench_r01_n_2gpu.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 | cut -c1-200
