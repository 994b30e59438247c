#!/bin/bash
# Round 2, GPU session A (1 GPU): build check, smoke, the whole -m gpu suite (no -x: list every failure), bench lines.
mkdir -p gpurun_out && rm -f gpurun_out/r02_*.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_gpu.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/a_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/a_smoke.log
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider --durations=15 > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
tail -40 gpurun_out/a_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --workload train --steps 20 --warmup 5 > gpurun_out/a_bench_train.json 2> gpurun_out/a_bench_train.err; echo "bench train rc=$?"
tail -5 gpurun_out/a_smoke.log; head -c 1500 gpurun_out/a_bench.json; echo; head -c 600 gpurun_out/a_bench_train.json; echo; tail -3 gpurun_out/a_bench.err
