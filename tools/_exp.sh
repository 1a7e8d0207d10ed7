#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do
timeout 600 python bench.py --frames 64 --steps 3 --warmup 3 --no-cpu-baseline --octomap-scans 32 > gpurun_out/bench_k3.json 2> gpurun_out/bench_k3.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_k3.json'))['octomap']
print('scans/s',round(d['value']),'ms/scan',round(d['ms_per_scan'],3),'kernel ms',round(d['raycast_kernel_ms_last_scan'],3),'un G pts/s', round(d['update_node']['value']/1e9,2), 'un ms', round(d['update_node']['ms'],2))
PY
done
nproc; uptime
