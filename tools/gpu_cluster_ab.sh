#!/bin/bash
# k_tc_pass as CTA pairs sharing one multicast weight stream (NSR_TC_CLUSTER=2): correctness, then A/B timing
mkdir -p gpurun_out
NSR_TC_CLUSTER=2 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_train.py tests/test_gpu_frame_parity.py -x -q -m gpu -p no:cacheprovider > gpurun_out/cl_pytest.log 2>&1; echo "pytest(cluster=2) rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed|nsr_tc\]|E   " gpurun_out/cl_pytest.log | tail -12
for C in 1 2 1 2; do echo "cluster=$C"; NSR_TC_CLUSTER=$C timeout 300 python tools/l2_weight_ab.py | grep '"flags": 0'; done
