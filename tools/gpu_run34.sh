timeout 300 python tools/gang_probe.py 4096 1,2,4,8,16,32,64,96 2>&1 | tail -9
timeout 120 python tools/gang_probe.py 1024 1,2,4,7,21,40,64,96 2>&1 | tail -9
timeout 120 python tools/gang_probe.py 2048 1,4,16,21,48,96 2>&1 | tail -7
