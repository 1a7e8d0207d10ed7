#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --frames 256 --steps 3 --warmup 3 --octomap-scans 0 > gpurun_out/bench_text.json 2> gpurun_out/bench_text.err; echo "bench exit $?"; tail -3 gpurun_out/bench_text.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_text.json'))
print(json.dumps(d['text_rows'])[:1800])
PY
