#!/bin/bash
# N-GPU job: NCCL merge parity, bench under torchrun (fixed OctoMap workload), PCIe probe
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tests/workers/nccl_octomap_worker.py 19 3 > gpurun_out/mg${N}_worker.log 2>&1; echo "nccl worker exit $?"; tail -3 gpurun_out/mg${N}_worker.log
timeout 300 $TR --master-port 29533 tools/pcie_probe.py > gpurun_out/mg${N}_pcie.json 2> gpurun_out/mg${N}_pcie.err; echo "pcie exit $?"; cat gpurun_out/mg${N}_pcie.json
for ov in 1; do
R3D_MERGE_OVERLAP=$ov TORCH_NCCL_HIGH_PRIORITY=1 timeout 900 $TR --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 --frames 2048 --quick > gpurun_out/mg${N}_bench_ov$ov.json 2> gpurun_out/mg${N}_bench_ov$ov.err; echo "bench ov=$ov exit $?"
python - $N $ov <<'PY'
import json,sys
try:
    d=json.loads([l for l in open('gpurun_out/mg%s_bench_ov%s.json'%(sys.argv[1],sys.argv[2])).read().splitlines() if l.startswith('{')][-1])
    o=d['octomap']
    print('N',d['n_gpus'],'points',round(d['value']/1e9,1),'G/s e2e',round(d['e2e']['value']/1e9,2),'octomap scans/s',round(o['value']),'runs',[round(x,3) for x in o['ms_per_scan_runs']],'gather',[round(x,4) for x in o['brick_gather_s_runs']],o['bt_sha256'][:12],o['bt_identical_on_all_ranks'])
except Exception as e:
    print('failed',e)
PY
tail -3 gpurun_out/mg${N}_bench_ov$ov.err
done
