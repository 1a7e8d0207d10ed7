#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --frames 64 --steps 3 --warmup 3 --octomap-scans 16 > gpurun_out/bench_un.json 2> gpurun_out/bench_un.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_un.json'))
print(json.dumps(d['octomap']['update_node']))
PY
tail -3 gpurun_out/bench_un.err
CMD="python bench.py --frames 64 --steps 2 --warmup 3 --no-cpu-baseline --octomap-scans 12"
timeout 600 $CMD > gpurun_out/plain_un.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_points -c 40 --csv --log-file gpurun_out/launches_un.csv $CMD > gpurun_out/ncu_un.log 2>&1
grep -E "k_points" gpurun_out/launches_un.csv | awk -F'","' '{print $5, $NF}' | tail -12
