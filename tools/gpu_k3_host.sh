#!/bin/bash
# OctoMap section of the bench with the host-side pipeline clock (is the host or the GPU the slower side?)
mkdir -p gpurun_out
if [ "$1" = "--tests" ]; then timeout 600 python -m pytest tests/test_octree_gpu.py tests/test_multigpu_gpu.py -m gpu -q -x 2>&1 | tail -2; fi
for ov in 1 0; do
R3D_PIPE_OVERLAP=$ov timeout 600 python bench.py --frames 64 --steps 3 --warmup 3 --no-cpu-baseline --octomap-scans 32 > gpurun_out/bench_k3_ov$ov.json 2> gpurun_out/bench_k3.err
python - $ov <<'PY'
import json,sys
d=json.load(open('gpurun_out/bench_k3_ov%s.json' % sys.argv[1]))['octomap']
print('overlap',sys.argv[1],'scans/s',round(d['value']),'ms/scan runs',[round(x,3) for x in d['ms_per_scan_runs']],'kernel ms',round(d['raycast_kernel_ms_last_scan'],3))
for h in d['host_pipeline_runs'][:1]: print('   host', {k:(round(v,3) if isinstance(v,float) else v) for k,v in h.items()})
PY
done
