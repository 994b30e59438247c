#!/bin/bash
# dW experiment: training tests, step time, DRAM bytes of the two k_tg_dw launches of one step
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py -q -p no:cacheprovider -x > gpurun_out/dw_pytest.log 2>&1; tail -n 3 gpurun_out/dw_pytest.log
for i in 1 2 3; do timeout 120 python tools/train_bench.py 512 30; done 2>&1 | grep -o '"ms_per_step": [0-9.]*\|"backward": [0-9.]*'
timeout 120 python tools/train_bench.py 64 50 2>&1 | grep -o '"ms_per_step": [0-9.]*'
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none -k regex:k_tg_dw -s 4 -c 2 --csv --log-file gpurun_out/dw_ncu.csv python tools/train_bench.py 512 2 > /dev/null 2>&1
grep -v "^==" gpurun_out/dw_ncu.csv | cut -d, -f5,13- | cut -c1-200
