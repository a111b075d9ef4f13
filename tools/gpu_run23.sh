for f in 0 4; do
echo "S4 flags $f: $(PK_TN_FLAGS=$f PK_PROBE_TIME=1 PK_PROBE_REPS=5 timeout 300 python tools/predict_probe.py 2>&1 | grep predictions)"
done
cp pykriging_b200/libpykriging_b200.so /tmp/keep.so; cp pykriging_b200/lib_s5.so.bin pykriging_b200/libpykriging_b200.so; touch pykriging_b200/libpykriging_b200.so
for f in 0 1 4; do
echo "S5 flags $f: $(PK_TN_FLAGS=$f PK_PROBE_TIME=1 PK_PROBE_REPS=5 timeout 300 python tools/predict_probe.py 2>&1 | grep predictions)"
done
