#!/bin/bash
# Round 2: the one-launch frame kernel -- bit-equality tests first (under a short timeout: a protocol bug would hang),
# then the whole GPU suite (which now runs on the one-launch path wherever the option set allows), then small-size timings
# with and without it (NSR_TC_FUSED=0 keeps the separate launches).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fused_frame.py -q -p no:cacheprovider -x --timeout 120 > gpurun_out/f_fused.log 2>&1; echo "fused rc=$?" >> gpurun_out/f_fused.log
tail -25 gpurun_out/f_fused.log
if grep -q "fused rc=0" gpurun_out/f_fused.log; then
  timeout 600 python tools/config_sweep.py > gpurun_out/f_sweep.txt 2>&1; cat gpurun_out/f_sweep.txt
  NSR_TC_FUSED=0 timeout 600 python tools/config_sweep.py > gpurun_out/f_sweep_separate.txt 2>&1; cat gpurun_out/f_sweep_separate.txt
fi
