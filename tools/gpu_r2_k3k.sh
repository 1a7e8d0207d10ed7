#!/bin/bash
mkdir -p gpurun_out
run() { # label lib env
  env $3 R3D_LIB_PATH=$PWD/3d_reconstruction_system_b200/libr3d_b200$2.so timeout 600 python bench.py --frames 2048 --steps 3 --warmup 3 --no-cpu-baseline --quick --octomap-scans 1024 > gpurun_out/k3k_$1.json 2> gpurun_out/k3k_$1.err
  python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open('gpurun_out/k3k_%s.json'%sys.argv[1]))['octomap']
    print('%-12s'%sys.argv[1],'scans/s',round(d['value']),[round(x,3) for x in d['ms_per_scan_runs']],'kernel',round(d['raycast_kernel_ms_per_scan_last_batch'],3),d['bt_sha256'][:12],'bt_write',round(d['bt_write_s'],3))
except Exception as e:
    print(sys.argv[1],'failed',e)
PY
}
run b8_c256 "" "R3D_SCAN_BATCH=8"
run b8_c384 _c384 "R3D_SCAN_BATCH=8"
run b8_c512 _c512 "R3D_SCAN_BATCH=8"
run b8_c1024 _c1024 "R3D_SCAN_BATCH=8"
run b4_c384 _c384 "A=1"
run b6_c256 "" "R3D_SCAN_BATCH=6"
timeout 600 python -m pytest tests/test_octree_gpu.py -m gpu -q -x -k "bt or binary or known or prune or pool_growth or sequence" > gpurun_out/k3k_pytest.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/k3k_pytest.log
