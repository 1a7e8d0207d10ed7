#!/bin/bash
# K3 variants (refill threshold, scans per batch) + ncu capture of the walk kernel
mkdir -p gpurun_out
run() { # tag lib batch
  R3D_LIB_PATH=$PWD/3d_reconstruction_system_b200/libr3d_b200$2.so R3D_SCAN_BATCH=$3 timeout 600 python bench.py --frames 256 --steps 3 --warmup 3 --no-cpu-baseline --octomap-scans 128 > gpurun_out/k3b_$1.json 2> gpurun_out/k3b_$1.err
  python - $1 <<'PY'
import json,sys
try:
    d=json.load(open('gpurun_out/k3b_%s.json'%sys.argv[1]))['octomap']
    print('%-10s'%sys.argv[1],'scans/s',round(d['value']),'ms/scan runs',[round(x,3) for x in d['ms_per_scan_runs']],'kernel ms/scan',round(d['raycast_kernel_ms_last_scan'],3),'host',d['host_pipeline_runs'][-1])
except Exception as e:
    print(sys.argv[1],'failed',e)
PY
}
run rf4_b4 "" 4
run rf4_b8 "" 8
run rf4_b2 "" 2
run rf4_b1 "" 1
run rf2_b4 _rf2 4
run rf8_b4 _rf8 4
CMD="python bench.py --frames 64 --steps 2 --warmup 3 --no-cpu-baseline --octomap-scans 16"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_scan_walk -s 3 -c 1 -o gpurun_out/k3_walk_prof -f $CMD > gpurun_out/ncu_k3walk.log 2>&1; echo "ncu exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/k3b_launches.csv $CMD > gpurun_out/ncu_k3b_l.log 2>&1; echo "ncu launches exit $?"
