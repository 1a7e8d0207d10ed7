#!/bin/bash
mkdir -p gpurun_out
for tag in "A --no-cpu-baseline" "B --frames 64" "C --frames 1024 --no-cpu-baseline"; do
  set -- $tag; name=$1; shift
  timeout 900 python bench.py --steps 3 --warmup 3 "$@" > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  python - "$name" <<'PY'
import json,sys
d=json.load(open('gpurun_out/bench_%s.json'%sys.argv[1])); o=d['octomap']
print(sys.argv[1], 'octo',round(o['value']),'ms/scan',round(o['ms_per_scan'],3),'kernel',round(o['raycast_kernel_ms_last_scan'],3),'bt_write',round(o['bt_write_s'],3),'un_ms',round(o['update_node']['ms'],2))
PY
done
