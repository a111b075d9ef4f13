mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 tools/dist_check.py > gpurun_out/dist_check_4gpu_r01_s4.log 2>&1; echo "dist rc=$?"; grep -v "^\s*$" gpurun_out/dist_check_4gpu_r01_s4.log | tail -4
