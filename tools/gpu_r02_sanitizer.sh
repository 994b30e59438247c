#!/bin/bash
# compute-sanitizer memcheck over the kernels that changed in round 2 (forward as CTA pairs, the one-launch frame kernel,
# dX chain, persistent dW, render backward with the new tile images, the P2P all-reduce on one device)
mkdir -p gpurun_out
timeout 2400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest -q -m gpu -p no:cacheprovider -x \
  "tests/test_gpu_parity.py::test_forward_rays_against_reference_golden" \
  "tests/test_gpu_parity.py::test_tiny_and_ragged_batches" \
  "tests/test_gpu_fused_frame.py" \
  "tests/test_gpu_train.py::test_stash_matches_oracle_activations" \
  "tests/test_gpu_train.py::test_dx_gemm_against_torch" "tests/test_gpu_train.py::test_dw_gemm_against_torch" \
  "tests/test_gpu_train.py::test_fused_dx_chain_equals_layerwise_kernels" \
  "tests/test_gpu_train.py::test_trainer_trajectory_against_oracle" \
  "tests/test_gpu_comm.py" > gpurun_out/san_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -8 gpurun_out/san_memcheck.log
