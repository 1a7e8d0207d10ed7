#!/bin/bash
mkdir -p gpurun_out
for sp in 1 2 1 2; do
  R3D_STAGE_SPLIT=$sp timeout 200 python tools/e2e_probe.py 1500 2>&1 | tail -1
done
R3D_STAGE_SPLIT=2 R3D_STAGE_CHUNK_MB=512 timeout 200 python tools/e2e_probe.py 1500 2>&1 | tail -1
timeout 600 python bench.py --frames 256 --steps 3 --warmup 3 --octomap-scans 8 > gpurun_out/bench_text.json 2> gpurun_out/bench_text.err; echo "bench exit $?"; tail -3 gpurun_out/bench_text.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_text.json'))
print(json.dumps(d['text_rows'])[:1500])
print(json.dumps(d['compact_mode'])[:300])
print(d['octomap']['value'], d['octomap'].get('ms_per_scan_runs'), d['octomap']['update_node']['ms_runs'])
PY
