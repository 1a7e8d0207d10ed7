#!/bin/bash
# parity of the occupancy path + the OctoMap section of the bench on a short frame stack (K3 iteration loop)
mkdir -p gpurun_out
TAG=${1:-k3}
timeout 1500 python -m pytest tests/test_octree_gpu.py tests/test_multigpu_gpu.py tests/test_scripts_gpu.py -m gpu -q -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -12 gpurun_out/${TAG}_pytest.log
for scans in 32 128; do
timeout 600 python bench.py --frames 256 --steps 3 --warmup 3 --no-cpu-baseline --octomap-scans $scans > gpurun_out/${TAG}_bench$scans.json 2> gpurun_out/${TAG}_bench$scans.err; echo "bench exit $?"
python - $TAG $scans <<'PY'
import json,sys
try:
    d=json.load(open('gpurun_out/%s_bench%s.json'%(sys.argv[1],sys.argv[2])))['octomap']
    print('scans/s',round(d['value']),'ms/scan runs',[round(x,3) for x in d['ms_per_scan_runs']],'kernel ms/scan',round(d['raycast_kernel_ms_last_scan'],3),'steps/scan',d['dda_steps_per_scan'],'parity',d.get('parity_bt_ok'),'host',d['host_pipeline_runs'][-1])
except Exception as e:
    print('failed',e)
PY
tail -3 gpurun_out/${TAG}_bench$scans.err
done
