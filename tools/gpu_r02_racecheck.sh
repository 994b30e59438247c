#!/bin/bash
# compute-sanitizer racecheck (shared-memory hazards between the warp roles) over the one-launch frame kernel on small batches.
# racecheck does not model every mbarrier / async-proxy edge: read its report as a list of candidates, not as a verdict.
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 9 python -m pytest -q -p no:cacheprovider -x \
  "tests/test_gpu_fused_frame.py::test_forward_rays_in_one_launch_is_bit_identical" \
  "tests/test_gpu_fused_frame.py::test_box_average_in_the_compositing_epilogue_is_bit_identical" -k "bf16x3 or 75 or 37" > gpurun_out/san_racecheck.log 2>&1; echo "racecheck rc=$?"
grep -c "Race reported\|hazard" gpurun_out/san_racecheck.log; tail -n 25 gpurun_out/san_racecheck.log | cut -c1-300
