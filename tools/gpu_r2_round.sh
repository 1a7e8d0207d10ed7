#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu_gpu.py tests/test_text_gpu.py -m gpu -q -x > gpurun_out/round_pytest.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/round_pytest.log
bash tools/gpu_r2_mg.sh ${1:-2}
