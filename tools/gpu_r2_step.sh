#!/bin/bash
# 1-GPU check after the end-of-batch change: octree + multi-GPU-logic tests, two traced quick benches (pass-to-pass spread),
# then the ncu re-capture of the changed kernel families
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_octree_gpu.py tests/test_multigpu_gpu.py tests/test_scripts_gpu.py -m gpu -q -x > gpurun_out/step_pytest.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/step_pytest.log
for k in 1 2; do
R3D_PIPE_TRACE=1 timeout 600 python bench.py --frames 2048 --steps 3 --warmup 3 --no-cpu-baseline --quick --octomap-scans 1024 > gpurun_out/step_bench$k.json 2> gpurun_out/step_bench$k.err; grep "r3d pipe" gpurun_out/step_bench$k.err | grep "scans:" | tail -3
python - $k <<'PY'
import json,sys
d=json.load(open('gpurun_out/step_bench%s.json'%sys.argv[1]))['octomap']
print('scans/s',round(d['value']),[round(x,3) for x in d['ms_per_scan_runs']],'kernel',round(d['raycast_kernel_ms_per_scan_last_batch'],3),d['bt_sha256'][:12],'wall',[[round(a),round(b)] for a,b in d['host_wall_ms_runs_call_and_drain']])
PY
done
bash tools/gpu_r2_ncu2.sh
