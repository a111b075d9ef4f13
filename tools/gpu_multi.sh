# 8-GPU bench + sharding check (gpurun --gpus 8)
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_r01_s4_8gpu.json 2> gpurun_out/bench8_err.log; echo "bench8 rc=$?"; tail -c 400 gpurun_out/bench8_err.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r01_s4_8gpu.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['value'], d['e2e']['value'], d['ms_per_step'], d['predictions']['value'], d['clocks'])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 tools/dist_check.py > gpurun_out/dist_check_8gpu_r01_s4.log 2>&1; echo "dist rc=$?"; tail -4 gpurun_out/dist_check_8gpu_r01_s4.log
