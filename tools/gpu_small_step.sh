#!/bin/bash
# Is the 256-ray (DDP shape at 8 GPUs) training step host-bound?  Event-timed step vs the sum of its kernel durations.
mkdir -p gpurun_out
python tools/train_bench.py 64 50
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 94 --csv --log-file gpurun_out/s_launches.csv python tools/train_bench.py 64 4 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/s_launches.csv')) if len(r)>10 and r[0].isdigit()]
tot=sum(float(r[-1].replace(',','')) for r in rows)
print("kernel time of 94 launches (2 steps):", round(tot/1e3,1), "us ->", round(tot/2e3,1), "us per step")
PY
