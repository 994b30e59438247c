#!/bin/bash
# third short GPU call: training tests with the dW stream fan-out, A/B timing against the single-stream order
mkdir -p gpurun_out
T0=$(date +%s)
timeout 300 python -m pytest tests/test_gpu_train.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_train.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s) - T0 ))s" >> gpurun_out/pytest_train.log
for f in 0 4 0 4; do NSR_DEBUG_FLAGS=$f timeout 120 python tools/train_bench.py 512 30 >> gpurun_out/train_ab.jsonl 2>> gpurun_out/train_ab.err; done
NSR_DEBUG_FLAGS=0 timeout 120 python tools/train_bench.py 2048 10 >> gpurun_out/train_ab.jsonl 2>> gpurun_out/train_ab.err
NSR_DEBUG_FLAGS=4 timeout 120 python tools/train_bench.py 2048 10 >> gpurun_out/train_ab.jsonl 2>> gpurun_out/train_ab.err
timeout 120 python bench.py --workload train --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err
echo "t=$(( $(date +%s) - T0 ))s"
tail -5 gpurun_out/pytest_train.log; cat gpurun_out/train_ab.jsonl; tail -3 gpurun_out/train_ab.err; cat gpurun_out/bench_train.json | cut -c1-400
