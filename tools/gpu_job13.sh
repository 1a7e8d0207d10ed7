#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_k1_gpu.py tests/test_octree_gpu.py -m gpu -q -x --durations=6 -k "c5_shape or c2_full or c4_airsim" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
