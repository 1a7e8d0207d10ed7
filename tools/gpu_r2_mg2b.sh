#!/bin/bash
# 2-GPU check of the cast gate: NCCL merge parity worker, then the scaling command with the gate on and off
N=2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tests/workers/nccl_octomap_worker.py 19 3 > gpurun_out/mg${N}_worker.log 2>&1; echo "nccl worker exit $?"; tail -2 gpurun_out/mg${N}_worker.log
for gate in 1 0; do
R3D_CAST_GATE=$gate TORCH_NCCL_HIGH_PRIORITY=1 timeout 900 $TR --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 --frames 2048 --quick > gpurun_out/mg2_gate$gate.json 2> gpurun_out/mg2_gate$gate.err; echo "bench gate=$gate exit $?"
python - gpurun_out/mg2_gate$gate.json <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith('{')][-1])
    o=d['octomap']
    print('N',d['n_gpus'],'points',round(d['value']/1e9,1),'e2e',d['e2e'] and round(d['e2e']['value']/1e9,2),'octomap',round(o['value']),'scans/s runs',[round(x,4) for x in o['ms_per_scan_runs']],'gather',[round(x,4) for x in o.get('brick_gather_s_runs',[])],o['bt_sha256'][:12],o.get('bt_identical_on_all_ranks'),'bt_write',round(o['bt_write_s'],3))
except Exception as e:
    print('failed',e)
PY
done
cp gpurun_out/mg2_gate1.json gpurun_out/r2_bench_2gpu_frames2048.json
