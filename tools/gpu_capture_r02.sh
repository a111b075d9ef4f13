# Round 2: bench + launch list + ncu --set full captures of the INT8 Cholesky and the Psi kernel (run under gpurun;
# summaries -> profiles/ with tools/summarize_ncu.py)
mkdir -p gpurun_out
T=${1:-m}
timeout 900 python bench.py > gpurun_out/bench_r02_$T.json 2> gpurun_out/bench_err.log; echo "bench rc=$?"; tail -3 gpurun_out/bench_err.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02_$T.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$T.log 2>&1; tail -1 gpurun_out/ncu_$T.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:chol_i8 -c 1 -o gpurun_out/prof_chol_i8_r02_$T -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-predictions --no-extras > gpurun_out/ncu_${T}1.log 2>&1; tail -1 gpurun_out/ncu_${T}1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:psi_pop -c 1 -o gpurun_out/prof_psi_pop_r02_$T -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-predictions --no-extras > gpurun_out/ncu_${T}2.log 2>&1; tail -1 gpurun_out/ncu_${T}2.log
