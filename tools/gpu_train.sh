#!/bin/bash
# one GPU call for the training row: train tests (all, with diagnostics), the forward parity suite, a
# bench sanity line and the per-phase timing of the fused iteration
mkdir -p gpurun_out
rm -f gpurun_out/train_parity.jsonl
timeout 900 python -m pytest tests/test_gpu_train.py -q -p no:cacheprovider > gpurun_out/pytest_train.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_train.log
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
timeout 300 python tools/train_bench.py 512 20 > gpurun_out/train_bench.json 2> gpurun_out/train_bench.err
timeout 300 python tools/train_bench.py 2048 10 > gpurun_out/train_bench_8k.json 2>> gpurun_out/train_bench.err
tail -40 gpurun_out/pytest_train.log; tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_quick.json gpurun_out/train_bench.json gpurun_out/train_bench_8k.json; tail -5 gpurun_out/train_bench.err
