#!/bin/bash
# 1-GPU job: full GPU test suite, default bench, K3 occupancy experiments
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --durations=5 > gpurun_out/full1_pytest.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/full1_pytest.log
timeout 300 python tools/pcie_probe.py > gpurun_out/mg1_pcie.json 2> gpurun_out/mg1_pcie.err; cat gpurun_out/mg1_pcie.json
timeout 900 python bench.py --steps 5 > gpurun_out/full1_bench.json 2> gpurun_out/full1_bench.err; echo "bench exit $?"; tail -3 gpurun_out/full1_bench.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/full1_bench.json'))
    print('value',round(d['value']/1e9,1),'frac',round(d['roofline']['frac'],3),'e2e',round(d['e2e']['value']/1e9,2),d['e2e'].get('frac_of_measured_d2h_ceiling'))
    o=d['octomap']; print('octo',round(o['value']),[round(x,3) for x in o['ms_per_scan_runs']],o['raycast_kernel_ms_per_scan_last_batch'],o['parity_bt_ok'],o['bt_sha256'][:12],o['growth'])
    print('compact',d['compact_mode']); print('text',{k:(round(v['points_per_s']/1e9,2),round(v['frac_of_hbm_peak'],3),v.get('parity_ok')) for k,v in d['text_rows'].items() if isinstance(v,dict)})
    print('script',d['script_e2e']); print('png',d['png_decode'])
except Exception as e:
    print('parse failed',e)
PY
for tag in occ5 rf12; do
  R3D_LIB_PATH=$PWD/3d_reconstruction_system_b200/libr3d_b200_$tag.so timeout 600 python bench.py --frames 512 --steps 3 --warmup 3 --no-cpu-baseline --quick --octomap-scans 128 > gpurun_out/full1_k3_$tag.json 2> gpurun_out/full1_k3_$tag.err
  python - $tag <<'PY'
import json,sys
try:
    d=json.load(open('gpurun_out/full1_k3_%s.json'%sys.argv[1]))['octomap']
    print('%-8s'%sys.argv[1],'scans/s',round(d['value']),[round(x,3) for x in d['ms_per_scan_runs']],'kernel',round(d['raycast_kernel_ms_per_scan_last_batch'],3))
except Exception as e:
    print(sys.argv[1],'failed',e)
PY
done
