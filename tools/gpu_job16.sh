#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/sanitizer_case.py > gpurun_out/san_plain.log 2>&1; echo "plain exit $?"; tail -2 gpurun_out/san_plain.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitizer_case.py > gpurun_out/san_memcheck.log 2>&1; echo "memcheck exit $?"
tail -6 gpurun_out/san_memcheck.log
