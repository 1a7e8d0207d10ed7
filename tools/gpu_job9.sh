#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --frames 64 --steps 3 --warmup 3 --no-cpu-baseline --octomap-scans 32"
timeout 600 $CMD > gpurun_out/bench_k3.json 2> gpurun_out/bench_k3.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_k3.json'))
print(json.dumps(d['octomap']))
PY
tail -5 gpurun_out/bench_k3.err
