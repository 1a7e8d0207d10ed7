#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_scripts_gpu.py -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
