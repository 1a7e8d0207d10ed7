#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_text_gpu.py -m gpu -q -x > gpurun_out/k6_pytest.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/k6_pytest.log
for lib in "" $K6_VARIANTS; do
  R3D_LIB_PATH=$PWD/3d_reconstruction_system_b200/libr3d_b200$lib.so timeout 300 python tools/k6_probe.py 64 > gpurun_out/k6_probe$lib.json 2> gpurun_out/k6_probe$lib.err || tail -3 gpurun_out/k6_probe$lib.err
  python - "$lib" <<'PY'
import json,sys
try:
    d=json.load(open('gpurun_out/k6_probe%s.json'%sys.argv[1]))
    t=d['text_rows']
    print('%-6s'%(sys.argv[1] or 'deflt'),{k:(round(v['ms'],3),round(v['points_per_s']/1e9,2),round(v['frac_of_hbm_peak'],3),v.get('parity_ok')) for k,v in t.items() if isinstance(v,dict)})
except Exception as e:
    print(sys.argv[1],'failed',e)
PY
done
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:k6_rows -c 1 -o gpurun_out/r2_k6_rows_v3 -f python tools/k6_probe.py 64 > gpurun_out/ncu_k6.log 2>&1; echo "ncu exit $?"
R3D_PIPE_TRACE=1 timeout 600 python bench.py --frames 2048 --steps 3 --warmup 3 --no-cpu-baseline --quick --octomap-scans 512 > gpurun_out/k3j_bench.json 2> gpurun_out/k3j_bench.err; grep "r3d pipe" gpurun_out/k3j_bench.err | grep "scans:" | tail -5
python - <<'PY'
import json
d=json.load(open('gpurun_out/k3j_bench.json'))['octomap']
print('scans/s',round(d['value']),[round(x,3) for x in d['ms_per_scan_runs']],'kernel',round(d['raycast_kernel_ms_per_scan_last_batch'],3),d['bt_sha256'][:12],'wall',[[round(a),round(b)] for a,b in d['host_wall_ms_runs_call_and_drain']])
PY
