#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_octree_gpu.py tests/test_multigpu_gpu.py tests/test_scripts_gpu.py -m gpu -q -x --durations=4 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --frames 64 --steps 3 --warmup 3 --no-cpu-baseline --octomap-scans 32 > gpurun_out/bench_k3.json 2> gpurun_out/bench_k3.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_k3.json'))
print(json.dumps(d['octomap']))
PY
tail -3 gpurun_out/bench_k3.err
