#!/bin/bash
mkdir -p gpurun_out
show() {
python - $1 <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    o=d['octomap']
    print(sys.argv[1],'points',round(d['points']['value']/1e9,1),'G/s frac',round(d['roofline']['frac'],3),'octomap',round(o['value']),'ms/scan',round(o['ms_per_scan'],3),'growth',o.get('growth'),'host',o.get('host_pipeline'),o['bt_sha256'][:12])
except Exception as e:
    print(sys.argv[1],'parse failed',e)
PY
}
timeout 900 python -m pytest tests/test_octree_gpu.py -m gpu -q -x > gpurun_out/pool_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pool_pytest.log
R3D_PIPE_TRACE=1 timeout 900 python bench.py --config c3 --steps 3 --no-cpu-baseline > gpurun_out/pool_c3_vmm.json 2> gpurun_out/pool_c3_vmm.err; show gpurun_out/pool_c3_vmm.json; grep "r3d pipe" gpurun_out/pool_c3_vmm.err | sort -t' ' -k8 -n | tail -5
R3D_POOL_MALLOC=1 R3D_PIPE_TRACE=1 timeout 900 python bench.py --config c3 --steps 3 --no-cpu-baseline > gpurun_out/pool_c3_malloc.json 2> gpurun_out/pool_c3_malloc.err; show gpurun_out/pool_c3_malloc.json; grep "r3d pipe" gpurun_out/pool_c3_malloc.err | tail -4
timeout 900 python bench.py --config c4 --steps 3 > gpurun_out/r2_bench_c4.json 2> gpurun_out/pool_c4.err; show gpurun_out/r2_bench_c4.json
