#!/bin/bash
# A/B of the order of the three split-precision MMAs per k-step (NSR_MMA_ORDER builds; energy-bound kernel: does operand
# reuse between consecutive MMAs change the clock?)
mkdir -p gpurun_out
for rep in 1 2; do
for o in 0 1 2; do
  if [ $o = 0 ]; then unset NSR_LIB_PATH; else export NSR_LIB_PATH=$PWD/nerf_sr_b200/libnsr_b200_order$o.so; fi
  timeout 300 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline --no-torch-gpu-port 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('order $o', round(d['value']), round(d['ms_per_step'],2), d['clocks']['sm_mhz_in_kernel'], d['clocks']['power_w_median'])"
done; done
