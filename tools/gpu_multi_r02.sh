# 8-GPU bench (strong block: one population of 1024 sharded over the ranks) + the multi-GPU pytest cases (gpurun --gpus 8)
mkdir -p gpurun_out
T=${1:-n}
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_r02_${T}_8gpu.json 2> gpurun_out/bench8_err.log; echo "bench8 rc=$?"; tail -c 300 gpurun_out/bench8_err.log
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r02_${T}_8gpu.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['value'], d['e2e']['value'], d['ms_per_step'], d['predictions']['value'])
print(json.dumps(d['strong']))
PY
timeout 300 python -m pytest tests -m gpu -x -q -k "two_devices or shard" 2>&1 | tail -3
