#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 --frames 2048 --quick --no-cpu-baseline > gpurun_out/mg1_bench.json 2> gpurun_out/mg1_bench.err; echo "exit $?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/mg1_bench.json').read().splitlines() if l.startswith('{')][-1]); o=d['octomap']
print('N 1 points',round(d['value']/1e9,1),'e2e',round(d['e2e']['value']/1e9,2),'octomap',round(o['value']),[round(x,3) for x in o['ms_per_scan_runs']],o['scans'],o['bt_sha256'])
PY
