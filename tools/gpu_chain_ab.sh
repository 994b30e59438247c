#!/bin/bash
# per-kernel timing of the training step under ncu for several debug-flag settings (timing experiments, wrong results)
mkdir -p gpurun_out
for F in 0 8 16 24; do
  NSR_DEBUG_FLAGS=$F timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 140 --csv --log-file gpurun_out/ab_$F.csv python tools/train_bench.py 512 4 > /dev/null 2>&1
  python - <<PY
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/ab_$F.csv')) if len(r)>10 and r[0].isdigit()]
agg=collections.OrderedDict()
for r in rows:
    name=r[4].split('(')[0][-30:]; t=float(r[-1].replace(',',''))
    a=agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=t
print("flags=$F", {k:(v[0], round(v[1]/1e3/v[0],1)) for k,v in agg.items() if 'tg_' in k or 'tc_pass' in k})
PY
done
