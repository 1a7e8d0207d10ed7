#!/bin/bash
# 1-GPU wrap-up: full GPU test suite, smoke, reference arm, default bench, configs 3 and 4 again (in-place pool, Z table)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --durations=5 > gpurun_out/final_pytest.log 2>&1; echo "pytest exit $?"; tail -9 gpurun_out/final_pytest.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/final_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/final_smoke.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/final_ref.err; echo "ref exit $?"
timeout 1200 python bench.py > gpurun_out/r2_bench_c2_default.json 2> gpurun_out/final_bench.err; echo "bench exit $?"; tail -3 gpurun_out/final_bench.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2_bench_c2_default.json'))
    print('value',round(d['value']/1e9,1),'frac',round(d['roofline']['frac'],3),'e2e',round(d['e2e']['value']/1e9,2),d['e2e'].get('frac_of_measured_d2h_ceiling'),'launches',d['gpu_launches'],d['clocks'])
    o=d['octomap']; print('octo',round(o['value']),[round(x,3) for x in o['ms_per_scan_runs']],o['raycast_kernel_ms_per_scan_last_batch'],o['parity_bt_ok'],o['scans'],o['bt_sha256'][:12],o['growth'],'un',round(o['update_node']['value']/1e9,1),o['update_node'].get('parity_bt_ok'))
    print('compact',d['compact_mode']); print('text',{k:(round(v['points_per_s']/1e9,2),round(v['frac_of_hbm_peak'],3),v.get('parity_ok')) for k,v in d['text_rows'].items() if isinstance(v,dict)})
    s=d['script_e2e']; print('script',s.get('frames_per_s'),s.get('points_per_s'),s.get('error')); print('png',d['png_decode'])
    r=json.load(open('gpurun_out/r2_bench_reference_arm.json')); print('ref',r['value'],r['config'])
except Exception as e:
    print('parse failed',e)
PY
for c in c3 c4; do
  timeout 1500 python bench.py --config $c --steps 3 > gpurun_out/r2_bench_$c.json 2> gpurun_out/r2_bench_$c.err; echo "bench $c exit $?"; tail -2 gpurun_out/r2_bench_$c.err
  python - $c <<'PY'
import json,sys
try:
    d=json.load(open('gpurun_out/r2_bench_%s.json'%sys.argv[1]))
    o=d['octomap']
    print(sys.argv[1],'value',d['value'],'points',round(d['points']['value']/1e9,1),'G/s frac',round(d['roofline']['frac'],3),'octomap',round(o['value']),'ms/scan',round(o['ms_per_scan'],3),'growth',o.get('growth'),'host',o.get('host_pipeline'),o['bt_sha256'][:12])
except Exception as e:
    print(sys.argv[1],'parse failed',e)
PY
done
