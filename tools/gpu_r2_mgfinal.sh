#!/bin/bash
# N-GPU wrap-up: (N=2: NCCL merge parity worker first) the scaling command of the driver, bench.py --frames 2048 --quick
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$N" = "2" ]; then
  timeout 600 $TR --master-port 29511 tests/workers/nccl_octomap_worker.py 19 3 > gpurun_out/mg${N}_worker.log 2>&1; echo "nccl worker exit $?"; tail -2 gpurun_out/mg${N}_worker.log
fi
TORCH_NCCL_HIGH_PRIORITY=1 timeout 900 $TR --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 --frames 2048 --quick > gpurun_out/r2_bench_${N}gpu_frames2048.json 2> gpurun_out/r2_bench_${N}gpu_frames2048.err; echo "bench exit $?"
python - gpurun_out/r2_bench_${N}gpu_frames2048.json <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith('{')][-1])
    o=d['octomap']
    print('N',d['n_gpus'],d['scaling'],'points',round(d['value']/1e9,1),'G/s frac',round(d['roofline']['frac'],3),'e2e',d['e2e'] and round(d['e2e']['value']/1e9,2),'octomap',round(o['value']),'scans/s runs',[round(x,4) for x in o['ms_per_scan_runs']],'gather',[round(x,4) for x in o.get('brick_gather_s_runs',[])],o['scans'],o['bt_sha256'][:12],o.get('bt_identical_on_all_ranks'))
except Exception as e:
    print('failed',e)
PY
tail -3 gpurun_out/r2_bench_${N}gpu_frames2048.err
