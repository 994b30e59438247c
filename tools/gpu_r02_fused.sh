#!/bin/bash
# Round 2: the one-launch frame kernel -- bit-equality tests first (under a short timeout: a protocol bug would hang),
# then the whole GPU suite (which now runs on the one-launch path wherever the option set allows), then small-size timings.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fused_frame.py -q -p no:cacheprovider -x --timeout 120 > gpurun_out/f_fused.log 2>&1; echo "fused rc=$?" >> gpurun_out/f_fused.log
tail -25 gpurun_out/f_fused.log
if grep -q "fused rc=0" gpurun_out/f_fused.log; then
  timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f_pytest.log
  tail -15 gpurun_out/f_pytest.log
  timeout 600 python tools/config_sweep.py > gpurun_out/f_sweep.txt 2>&1; cat gpurun_out/f_sweep.txt
  NSR_TC_FUSED=0 timeout 600 python tools/config_sweep.py > gpurun_out/f_sweep_separate.txt 2>&1; cat gpurun_out/f_sweep_separate.txt
  timeout 600 python bench.py --steps 10 --warmup 3 --no-extras > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "bench rc=$?"; head -c 900 gpurun_out/f_bench.json; echo
fi
