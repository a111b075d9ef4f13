timeout 300 python tools/train_probe.py 1024 10 1024 2>&1 | tail -2
timeout 300 python tools/train_probe.py 300 10 300 2>&1 | tail -2
timeout 300 python tools/train_probe.py 40 3 256 2>&1 | tail -2
