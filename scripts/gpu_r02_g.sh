#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_projection_train_r02.csv python scripts/prof_projection_train.py 3 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_projection_train_r02.csv')) if len(r)>5]
hdr=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hdr]; kn=h.index('Kernel Name'); mv=h.index('Metric Value')
data=rows[hdr+1:]
third=len(data)//3
for r in data[2*third:]:
    print(f"{float(r[mv])/1e3:9.1f} us  {r[kn][:110]}")
PY
