mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "predict or large_n" 2>&1 | tail -8
PK_PROBE_TIME=1 timeout 300 python tools/predict_probe.py 2>&1 | tail -3
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_predict_split.csv python tools/predict_probe.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_predict_split.csv')) if len(r)>5]
hdr=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hdr]; kn=h.index('Kernel Name'); mv=h.index('Metric Value'); mu=h.index('Metric Unit')
for r in rows[hdr+1:][:6]:
    if 'psi_rows' in r[kn] or 'trinorm' in r[kn]: print(r[kn][:40], r[mv], r[mu])
PY
