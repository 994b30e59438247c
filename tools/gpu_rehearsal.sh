#!/bin/bash
# dress rehearsal of the driver's round-end sequence on one fresh box: pytest -m gpu -x, smoke(), bench (reference arm first)
mkdir -p gpurun_out
T0=$(date +%s)
timeout 300 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/rehearsal_pytest.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s) - T0 ))s" >> gpurun_out/rehearsal_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/rehearsal_smoke.log 2>&1; echo "smoke rc=$? t=$(( $(date +%s) - T0 ))s" >> gpurun_out/rehearsal_smoke.log
timeout 200 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/rehearsal_bench_ref.json 2> gpurun_out/rehearsal_bench.err; echo "ref rc=$? t=$(( $(date +%s) - T0 ))s" >> gpurun_out/rehearsal_bench.err
timeout 200 python bench.py --gpus 1 > gpurun_out/rehearsal_bench.json 2>> gpurun_out/rehearsal_bench.err; echo "ours rc=$? t=$(( $(date +%s) - T0 ))s" >> gpurun_out/rehearsal_bench.err
tail -3 gpurun_out/rehearsal_pytest.log; tail -5 gpurun_out/rehearsal_smoke.log; cat gpurun_out/rehearsal_bench_ref.json | cut -c1-700; cat gpurun_out/rehearsal_bench.json | cut -c1-300; grep "rc=" gpurun_out/rehearsal_bench.err
