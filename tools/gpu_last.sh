#!/bin/bash
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_scenes.py -x -q -m gpu -p no:cacheprovider > gpurun_out/last_scenes.log 2>&1; echo "rc=$?" >> gpurun_out/last_scenes.log
tail -15 gpurun_out/last_scenes.log
