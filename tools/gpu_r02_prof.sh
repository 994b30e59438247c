#!/bin/bash
# Round 2 profiling session (1 GPU): launch lists (shares) + ncu --set full of the dominant kernels.  Numbers taken under
# ncu are not bench values.
mkdir -p gpurun_out && rm -f gpurun_out/p_*.ncu-rep
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/p_launches_render.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --no-torch-gpu-port > gpurun_out/p_ncu_render.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/p_launches_train.csv python tools/train_bench.py 512 4 > gpurun_out/p_ncu_train.log 2>&1
# the one-launch frame kernel (second of two frames), then the same frame as separate coarse and fine launches
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tc_pass -s 1 -c 3 -f -o gpurun_out/p_tc python tools/ncu_target.py > gpurun_out/p_ncu_full_tc.log 2>&1
# training: forward with stash (coarse, fine), dX chain (fine, coarse) of the 4th step
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_tg_dxchain|k_tc_pass" -s 12 -c 4 -f -o gpurun_out/p_train python tools/train_bench.py 512 2 > gpurun_out/p_ncu_full_train.log 2>&1
# training: the persistent dW kernel (one launch per net) of the 3rd step
timeout 900 ncu --set full --clock-control none -k regex:k_tg_dw -s 4 -c 2 -f -o gpurun_out/p_dw python tools/train_bench.py 512 2 > gpurun_out/p_ncu_full_dw.log 2>&1
ls -la gpurun_out/*.ncu-rep
for f in gpurun_out/p_ncu_full_tc.log gpurun_out/p_ncu_full_train.log gpurun_out/p_ncu_full_dw.log; do tail -n 2 $f; done
