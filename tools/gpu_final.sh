#!/bin/bash
# one short GPU call: the tests added after the last GPU session first, then the bench lines, then the rest of the suite
mkdir -p gpurun_out
rm -f gpurun_out/train_parity.jsonl
T0=$(date +%s)
timeout 300 python -m pytest tests/test_gpu_scenes.py tests/test_gpu_train.py -m gpu -q -p no:cacheprovider \
    -k "scene or loss_epilogue or gradients_against_oracle or all_loss_terms" > gpurun_out/pytest_new.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s) - T0 ))s" >> gpurun_out/pytest_new.log
timeout 200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$? t=$(( $(date +%s) - T0 ))s" >> gpurun_out/bench.err
timeout 120 python bench.py --workload train --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_train.json 2>> gpurun_out/bench.err; echo "rc=$? t=$(( $(date +%s) - T0 ))s" >> gpurun_out/bench.err
timeout 420 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=15 \
    -k "not (scene or loss_epilogue or gradients_against_oracle or all_loss_terms)" > gpurun_out/pytest_rest.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s) - T0 ))s" >> gpurun_out/pytest_rest.log
tail -25 gpurun_out/pytest_new.log; cat gpurun_out/bench.json gpurun_out/bench_train.json; tail -3 gpurun_out/bench.err; tail -30 gpurun_out/pytest_rest.log
