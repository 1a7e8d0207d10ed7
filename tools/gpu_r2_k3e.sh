#!/bin/bash
mkdir -p gpurun_out
run() { # tag lib batch
  R3D_LIB_PATH=$PWD/3d_reconstruction_system_b200/libr3d_b200$2.so R3D_SCAN_BATCH=$3 timeout 600 python bench.py --frames 512 --steps 3 --warmup 3 --no-cpu-baseline --quick --octomap-scans 128 > gpurun_out/k3e_$1.json 2> gpurun_out/k3e_$1.err
  python - $1 <<'PY'
import json,sys
try:
    d=json.load(open('gpurun_out/k3e_%s.json'%sys.argv[1]))['octomap']
    print('%-12s'%sys.argv[1],'scans/s',round(d['value']),'ms/scan runs',[round(x,3) for x in d['ms_per_scan_runs']],'kernel ms/scan',round(d['raycast_kernel_ms_per_scan_last_batch'],3),'host',[(round(h['wait_ms']),round(h['work_ms']),round(h['max_turnaround_ms'],1)) for h in d['host_pipeline_runs']], d['bt_sha256'][:12], d['growth'])
except Exception as e:
    print(sys.argv[1],'failed',e)
PY
  tail -2 gpurun_out/k3e_$1.err
}
run st3_b4 "" 4
run blind_b4 _blind 4
run blindrf4_b4 _blindrf4 4
run blind_b8 _blind 8
R3D_LIB_PATH=$PWD/3d_reconstruction_system_b200/libr3d_b200_blind.so timeout 900 python -m pytest tests/test_octree_gpu.py -m gpu -q -x > gpurun_out/k3e_pytest.log 2>&1; echo "pytest(blind) exit $?"; tail -3 gpurun_out/k3e_pytest.log
CMD="python bench.py --frames 64 --steps 2 --warmup 3 --no-cpu-baseline --quick --octomap-scans 16"
R3D_LIB_PATH=$PWD/3d_reconstruction_system_b200/libr3d_b200_blind.so timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_scan_walk -s 3 -c 1 -o gpurun_out/k3_walk_prof_blind -f $CMD > gpurun_out/ncu_k3walk3.log 2>&1; echo "ncu exit $?"
