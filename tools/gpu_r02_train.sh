#!/bin/bash
# training-path tests (stop at the first failure), then the training bench line
mkdir -p gpurun_out && rm -f gpurun_out/train_parity.jsonl
timeout 900 python -m pytest tests/test_gpu_train.py -x -q -m gpu -p no:cacheprovider > gpurun_out/tr_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tr_pytest.log
grep -E "^(FAILED|ERROR)|passed|failed|rc=|nsr_tc\]|nsr_train\]|E   " gpurun_out/tr_pytest.log | tail -25
timeout 300 python bench.py --workload train --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/tr_bench_train.json 2> gpurun_out/tr_bench_train.err; echo "bench train rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/tr_bench_train.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks'])"
tail -3 gpurun_out/tr_bench_train.err
