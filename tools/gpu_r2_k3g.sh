#!/bin/bash
mkdir -p gpurun_out
for lib in _ch128 _ch192 _ch256 _ch512 _ch128 _ch256; do
  R3D_LIB_PATH=$PWD/3d_reconstruction_system_b200/libr3d_b200$lib.so timeout 600 python bench.py --frames 2048 --steps 3 --warmup 3 --no-cpu-baseline --quick --octomap-scans 512 > gpurun_out/k3g_bench$lib.json 2> gpurun_out/k3g_bench$lib.err
  python - "$lib" <<'PY'
import json,sys
try:
    d=json.load(open('gpurun_out/k3g_bench%s.json'%sys.argv[1]))['octomap']
    print('%-8s'%(sys.argv[1] or 'default'),'scans/s',round(d['value']),[round(x,3) for x in d['ms_per_scan_runs']],'kernel',round(d['raycast_kernel_ms_per_scan_last_batch'],3),d['bt_sha256'][:12],'wall(call,drain)',[[round(a),round(b)] for a,b in d['host_wall_ms_runs_call_and_drain']],'host',[(round(h['wait_ms']),round(h['work_ms'],1)) for h in d['host_pipeline_runs']])
except Exception as e:
    print(sys.argv[1],'failed',e)
PY
done
