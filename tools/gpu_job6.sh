#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_octree_gpu.py tests/test_multigpu_gpu.py -m gpu -q -x --durations=12 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
CMD="python bench.py --frames 64 --steps 3 --warmup 3 --no-cpu-baseline --octomap-scans 32"
timeout 600 $CMD > gpurun_out/bench_k3.json 2> gpurun_out/bench_k3.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_k3.json'))
print(json.dumps(d['octomap']))
PY
tail -5 gpurun_out/bench_k3.err
CMD="python bench.py --frames 64 --steps 2 --warmup 3 --no-cpu-baseline --octomap-scans 6"
timeout 600 $CMD > gpurun_out/plain_k3v2.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_scan_raycast -s 4 -c 1 -o gpurun_out/k3_prof_v2 -f $CMD > gpurun_out/ncu_k3v2.log 2>&1
tail -3 gpurun_out/ncu_k3v2.log
