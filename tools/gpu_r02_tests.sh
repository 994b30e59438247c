#!/bin/bash
# -m gpu suite only (no -x), reports into gpurun_out/
mkdir -p gpurun_out && rm -f gpurun_out/r02_*.jsonl gpurun_out/train_parity.jsonl
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider "$@" > gpurun_out/t_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_pytest.log
grep -E "^(FAILED|ERROR)|passed|failed|rc=" gpurun_out/t_pytest.log | tail -40
