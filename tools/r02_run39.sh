#!/bin/bash
mkdir -p gpurun_out
GTTS_PROFILE_TRAIN=1 timeout -k 10 300 python tools/gpu_diag.py profile_vjp > gpurun_out/pt.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/pt.log; exit 1; }
GTTS_PROFILE_TRAIN=1 timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"wgrad|col_sums|rows_reduce" --launch-skip 600 --launch-count 400 --csv --log-file gpurun_out/r02_ncu_wgrad.csv python tools/gpu_diag.py profile_vjp > gpurun_out/ncu_wg.log 2>&1; echo "ncu rc $?"
python - <<'PY'
import csv, collections, re
rows=[r for r in csv.reader(open('gpurun_out/r02_ncu_wgrad.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows[1:]:
    n=re.sub(r'\(.*','',r[ki])[:60]; agg[n][0]+=1; agg[n][1]+=float(r[vi].replace(',',''))
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1]): print(f"{v[1]/1e3:9.1f} us {v[0]:4d} x {v[1]/v[0]/1e3:7.2f} us  {k}")
PY
