timeout 600 python -m pytest tests -m gpu -x -q -k "gang_size or cholesky_variants or fit or n4096 or full_size_c3" 2>&1 | tail -3
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-predictions 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['roofline']['frac'], d['roofline']['phase_ms_per_step'])"
timeout 120 python tools/gang_probe.py 1024 1,21,96 2>&1 | tail -3
timeout 120 python tools/gang_probe.py 4096 1,16 2>&1 | tail -2
