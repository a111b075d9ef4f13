run() { echo "$1 :: $(env $1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-predictions 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['ms_per_step'],2))")"; }
run "PK_NONE=1"
run "PK_CHOL_CTAS_PER_SM=1"
run "PK_CHOL_CTAS_PER_SM=1 PK_SPLIT=148,148,148,148,148,148,136"
run "PK_CHOL_CTAS_PER_SM=1 PK_SPLIT=256,256,256,256"
run "PK_CHOL_CTAS_PER_SM=1 PK_SPLIT=512,512"
run "PK_SPLIT=296,296,296,136"
