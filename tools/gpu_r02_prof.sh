#!/bin/bash
# Round 2 profiling session (1 GPU): launch lists (shares) + ncu --set full of the dominant kernels.  Numbers taken under
# ncu are not bench values.
mkdir -p gpurun_out && rm -f gpurun_out/p_*.ncu-rep
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/p_launches_render.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --no-torch-gpu-port > gpurun_out/p_ncu_render.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 141 --csv --log-file gpurun_out/p_launches_train.csv python tools/train_bench.py 512 4 > gpurun_out/p_ncu_train.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tc_pass -s 2 -c 2 -f -o gpurun_out/p_tc python tools/ncu_target.py > gpurun_out/p_ncu_full_tc.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_tg_dxchain|k_tc_pass" -s 12 -c 4 -f -o gpurun_out/p_train python tools/train_bench.py 512 2 > gpurun_out/p_ncu_full_train.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:k_tg_dw -s 26 -c 13 -f -o gpurun_out/p_dw python tools/train_bench.py 512 2 > gpurun_out/p_ncu_full_dw.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -2 gpurun_out/p_ncu_full_tc.log gpurun_out/p_ncu_full_train.log gpurun_out/p_ncu_full_dw.log
