#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu_gpu.py -m gpu -q -x -k nccl > gpurun_out/pytest_mg.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_mg.log
tail -5 gpurun_out/pytest_mg.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 --frames 512 --octomap-scans 32 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench2 exit $?"
python - <<'PY'
import json
for ln in open('gpurun_out/bench_2gpu.json'):
    if ln.startswith('{'): print(json.dumps(json.loads(ln)['octomap']))
PY
tail -3 gpurun_out/bench_2gpu.err
