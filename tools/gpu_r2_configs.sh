#!/bin/bash
# BASELINE configs 3, 4, 5 on one GPU (committed as profiles/r2_bench_c{3,4,5}.json) + the reference arm of each
mkdir -p gpurun_out
for c in c3 c4 c5; do
  timeout 1500 python bench.py --config $c --steps 3 > gpurun_out/r2_bench_$c.json 2> gpurun_out/r2_bench_$c.err; echo "bench $c exit $?"; tail -2 gpurun_out/r2_bench_$c.err
  python - $c <<'PY'
import json,sys
try:
    d=json.load(open('gpurun_out/r2_bench_%s.json'%sys.argv[1]))
    o=d['octomap']
    print(sys.argv[1],d['metric'],'value',d['value'],'points',round(d['points']['value']/1e9,1),'G/s frac',round(d['roofline']['frac'],3),'e2e',d['e2e'] and round(d['e2e']['value']/1e9,2))
    print('   octomap',round(o['value']),'scans/s ms/scan',round(o['ms_per_scan'],3),'steps/scan',o.get('dda_steps_per_scan'),'growth',o.get('growth'),'bricks',o.get('bricks'),'bt',o['bt_bytes'],o['bt_sha256'][:12],'host',o.get('host_pipeline'),'cpu',o.get('cpu_baseline'))
except Exception as e:
    print(sys.argv[1],'parse failed',e)
PY
done
timeout 600 python bench.py --config c4 --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_c4_ref.json 2>/dev/null; head -c 600 gpurun_out/r2_bench_c4_ref.json; echo
