#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-step2}
R3D_LIB_PATH=$PWD/3d_reconstruction_system_b200/libr3d_b200_$TAG.so timeout 900 python -m pytest tests/test_octree_gpu.py tests/test_multigpu_gpu.py -m gpu -q -x > gpurun_out/k3f_pytest.log 2>&1; echo "pytest($TAG) exit $?"; tail -3 gpurun_out/k3f_pytest.log
for lib in "" _$TAG; do
  R3D_LIB_PATH=$PWD/3d_reconstruction_system_b200/libr3d_b200$lib.so timeout 600 python bench.py --frames 512 --steps 3 --warmup 3 --no-cpu-baseline --quick --octomap-scans 256 > gpurun_out/k3f_bench$lib.json 2> gpurun_out/k3f_bench$lib.err
  python - "$lib" <<'PY'
import json,sys
try:
    d=json.load(open('gpurun_out/k3f_bench%s.json'%sys.argv[1]))['octomap']
    print('%-8s'%(sys.argv[1] or 'default'),'scans/s',round(d['value']),[round(x,3) for x in d['ms_per_scan_runs']],'kernel',round(d['raycast_kernel_ms_per_scan_last_batch'],3),d['bt_sha256'][:12])
except Exception as e:
    print(sys.argv[1],'failed',e)
PY
done
CMD="python bench.py --frames 64 --steps 2 --warmup 3 --no-cpu-baseline --quick --octomap-scans 16"
R3D_LIB_PATH=$PWD/3d_reconstruction_system_b200/libr3d_b200_$TAG.so timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_scan_walk -s 3 -c 1 -o gpurun_out/k3_walk_prof_$TAG -f $CMD > gpurun_out/ncu_k3f.log 2>&1; echo "ncu exit $?"
