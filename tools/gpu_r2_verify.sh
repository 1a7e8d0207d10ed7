#!/bin/bash
# last check of the round: the whole GPU suite and smoke() on the final tree
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/verify_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/verify_pytest.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/verify_smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/verify_smoke.log
