#!/bin/bash
mkdir -p gpurun_out
run() { # tag lib batch
  R3D_PIPE_TRACE=1 R3D_LIB_PATH=$PWD/3d_reconstruction_system_b200/libr3d_b200$2.so R3D_SCAN_BATCH=$3 timeout 600 python bench.py --frames 256 --steps 3 --warmup 3 --no-cpu-baseline --octomap-scans 128 > gpurun_out/k3c_$1.json 2> gpurun_out/k3c_$1.err
  python - $1 <<'PY'
import json,sys
try:
    d=json.load(open('gpurun_out/k3c_%s.json'%sys.argv[1]))['octomap']
    print('%-10s'%sys.argv[1],'scans/s',round(d['value']),'ms/scan runs',[round(x,3) for x in d['ms_per_scan_runs']],'kernel ms/scan',round(d['raycast_kernel_ms_last_scan'],3),'host',[(round(h['wait_ms']),round(h['work_ms']),round(h['max_turnaround_ms'],1)) for h in d['host_pipeline_runs']])
except Exception as e:
    print(sys.argv[1],'failed',e)
PY
  grep "r3d pipe" gpurun_out/k3c_$1.err | head -12
}
timeout 900 python -m pytest tests/test_octree_gpu.py tests/test_k1_gpu.py -m gpu -q -x > gpurun_out/k3c_pytest.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/k3c_pytest.log
run st3_b4 "" 4
run st2_b4 _st2 4
run st4_b4 _st4 4
run st3rf4_b4 _st3rf4 4
run st3ch32_b4 _st3ch32 4
run st3_b8 "" 8
CMD="python bench.py --frames 64 --steps 2 --warmup 3 --no-cpu-baseline --octomap-scans 16"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_scan_walk -s 3 -c 1 -o gpurun_out/k3_walk_prof2 -f $CMD > gpurun_out/ncu_k3walk2.log 2>&1; echo "ncu exit $?"
