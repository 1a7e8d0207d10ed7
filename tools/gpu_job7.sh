#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_octree_gpu.py tests/test_multigpu_gpu.py -m gpu -q -x --durations=5 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
CMD="python bench.py --frames 64 --steps 3 --warmup 3 --no-cpu-baseline --octomap-scans 32"
timeout 600 $CMD > gpurun_out/bench_k3.json 2> gpurun_out/bench_k3.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_k3.json'))
print(json.dumps(d['octomap']))
PY
tail -5 gpurun_out/bench_k3.err
