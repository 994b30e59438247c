#!/bin/bash
# first GPU bring-up: per-precision diagnostics, each in its own process under a timeout
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for p in fp32_simt bf16x3 fp16x3; do
  timeout 300 python tests/diag_parity.py $p > gpurun_out/diag_$p.log 2>&1
  echo "exit $p: $?" >> gpurun_out/diag_$p.log
done
tail -n 40 gpurun_out/diag_*.log
