mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trinorm -c 1 -s 2 -o gpurun_out/prof_trinorm_r01_a -f python tools/predict_probe.py > gpurun_out/ncu_tn.log 2>&1; tail -2 gpurun_out/ncu_tn.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:psi_rows -c 1 -s 2 -o gpurun_out/prof_psi_rows_r01_a -f python tools/predict_probe.py > gpurun_out/ncu_pr.log 2>&1; tail -2 gpurun_out/ncu_pr.log
