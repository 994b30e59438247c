#!/bin/bash
# ncu launch list (per-kernel durations) of the fused training iteration
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 140 --csv --log-file gpurun_out/train_launches.csv python tools/train_bench.py ${1:-2048} 4 > gpurun_out/ncu_train.log 2>&1
tail -3 gpurun_out/ncu_train.log
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/train_launches.csv')) if len(r)>10 and r[0].isdigit()]
agg=collections.OrderedDict()
for r in rows:
    name=r[4].split('(')[0][-40:]; t=float(r[-1].replace(',',''))
    a=agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=t
tot=sum(v[1] for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1]): print(f"{k:42s} n={v[0]:4d} total={v[1]/1e3:10.1f} us  {100*v[1]/tot:5.1f}%")
PY
