mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-predictions 2>gpurun_out/err.log | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['phase_ms_per_step'])" || tail -5 gpurun_out/err.log; }
run PK_PSI_CPI=1 PK_PSI_TAB=4096
run PK_PSI_CPI=2 PK_PSI_TAB=4096
PK_PSI_CPI=1 PK_PSI_TAB=4096 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
