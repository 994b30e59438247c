#!/bin/bash
# second short GPU call of the round's last session: full GPU suite, default bench line, memcheck over the new kernels
mkdir -p gpurun_out
T0=$(date +%s)
timeout 300 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s) - T0 ))s" >> gpurun_out/pytest_gpu.log
timeout 200 python bench.py --steps 5 --warmup 3 --torch-gpu-port > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$? t=$(( $(date +%s) - T0 ))s" >> gpurun_out/bench.err
timeout 240 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_scenes.py tests/test_gpu_train.py -m gpu -q -p no:cacheprovider \
    -k "scene or loss_epilogue or checkpoint" > gpurun_out/memcheck_new.log 2>&1; echo "memcheck rc=$? t=$(( $(date +%s) - T0 ))s" >> gpurun_out/memcheck_new.log
tail -6 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.json; tail -2 gpurun_out/bench.err; tail -12 gpurun_out/memcheck_new.log
