#!/bin/bash
# k_grad_reduce experiment: training tests, step time, per-launch duration of the reduce under ncu
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py tests/test_gpu_comm.py -q -p no:cacheprovider -x > gpurun_out/red_pytest.log 2>&1; tail -n 2 gpurun_out/red_pytest.log
for i in 1 2 3; do timeout 120 python tools/train_bench.py 512 30; done 2>&1 | grep -o '"ms_per_step": [0-9.]*'
timeout 120 python tools/train_bench.py 64 50 2>&1 | grep -o '"ms_per_step": [0-9.]*'
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k regex:k_grad_reduce -s 4 -c 2 --csv --log-file gpurun_out/red_ncu.csv python tools/train_bench.py 512 2 > /dev/null 2>&1
grep -v "^==" gpurun_out/red_ncu.csv | cut -d, -f5,13- | cut -c1-160
