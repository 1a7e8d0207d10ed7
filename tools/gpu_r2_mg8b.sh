#!/bin/bash
# 8-GPU job 2: scans-per-round 32 on the fixed workload, and BASELINE config 5 frame-sharded (points: all 10 000 frames with
# per-rank offsets; octree: the first 2 048 frames dealt out in rounds, merged)
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
show() {
python - $1 <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith('{')][-1])
    o=d['octomap']
    print(sys.argv[1],'N',d['n_gpus'],d['scaling'],'points',round(d['points']['value']/1e9,1),'G/s frac',round(d['roofline']['frac'],3),'e2e',d['e2e'] and round(d['e2e']['value']/1e9,2),'octomap',round(o['value']),'scans/s runs',[round(x,3) for x in o['ms_per_scan_runs']],'gather',[round(x,4) for x in o.get('brick_gather_s_runs',[])],o['scans'],o['bt_sha256'][:12],o.get('bt_identical_on_all_ranks'))
except Exception as e:
    print(sys.argv[1],'failed',e)
PY
}
TORCH_NCCL_HIGH_PRIORITY=1 timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 --frames 2048 --quick --octomap-scans-per-round 32 > gpurun_out/mg${N}_bench_c32.json 2> gpurun_out/mg${N}_bench_c32.err; echo "bench C=32 exit $?"; show gpurun_out/mg${N}_bench_c32.json
TORCH_NCCL_HIGH_PRIORITY=1 timeout 900 $TR --master-port 29513 bench.py --gpus $N --config c5 --steps 3 --warmup 3 --octomap-scans-full 2048 > gpurun_out/r2_bench_c5_${N}gpu.json 2> gpurun_out/r2_bench_c5_${N}gpu.err; echo "bench c5 exit $?"; show gpurun_out/r2_bench_c5_${N}gpu.json; tail -3 gpurun_out/r2_bench_c5_${N}gpu.err
