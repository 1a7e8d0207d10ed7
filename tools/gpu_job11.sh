#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_octree_gpu.py -m gpu -q -x --durations=3 -k "read_binary or bt_known" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
