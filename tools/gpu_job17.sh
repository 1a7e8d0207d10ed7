#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_k1_gpu.py -m gpu -q -x --durations=3 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
timeout 600 python tools/compact_probe.py > gpurun_out/compact_probe.log 2>&1; echo "exit $?"; tail -3 gpurun_out/compact_probe.log
