#!/bin/bash
# parity of the occupancy path + the OctoMap section of the bench on a short frame stack
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_octree_gpu.py tests/test_multigpu_gpu.py -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --frames 64 --steps 3 --warmup 3 --no-cpu-baseline --octomap-scans 32 > gpurun_out/bench_k3.json 2> gpurun_out/bench_k3.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_k3.json'))['octomap']
print('scans/s',round(d['value']),'ms/scan',round(d['ms_per_scan'],3),'kernel ms',round(d['raycast_kernel_ms_last_scan'],3),'steps/s in kernel %.1f G'%(d['raycast_steps_per_s_in_kernel']/1e9))
PY
tail -3 gpurun_out/bench_k3.err
