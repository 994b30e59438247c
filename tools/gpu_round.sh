#!/bin/bash
# one GPU call: parity tests, bench lines, ncu launch list + full capture of the tcgen05 kernel
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?" >> gpurun_out/bench.err
timeout 300 python bench.py --steps 3 --precision fp16x3 --no-cpu-baseline > gpurun_out/bench_fp16x3.json 2>> gpurun_out/bench.err
timeout 300 python bench.py --steps 2 --precision fp32_simt --no-cpu-baseline > gpurun_out/bench_simt.json 2>> gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tc_pass -s 2 -c 2 -f -o gpurun_out/prof_tc python tools/ncu_target.py > gpurun_out/ncu_full.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.json gpurun_out/bench_fp16x3.json gpurun_out/bench_simt.json gpurun_out/bench_reference.json; tail -3 gpurun_out/bench.err; tail -3 gpurun_out/ncu_full.log
