mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-predictions 2>gpurun_out/err.log | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['phase_ms_per_step'])" || tail -5 gpurun_out/err.log; }
run PK_SPLIT=1024
run PK_X=default
run PK_SPLIT=432,592
run PK_SPLIT=136,296,592
run PK_SPLIT=136,888
run PK_SPLIT=296,136,592
run PK_SPLIT=128,128,128,128,128,128,128,128
run PK_SPLIT=256,256,256,256
run PK_SPLIT=512,512
run PK_SPLIT=728,296
run PK_SPLIT=64,72,296,592
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
