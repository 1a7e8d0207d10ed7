#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8 > gpurun_out/mg_smi.txt
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "bench$N exit $?"
tail -c 2500 gpurun_out/bench_${N}gpu.json; tail -3 gpurun_out/bench_${N}gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_${N}gpu.json 2> gpurun_out/bench_ref_${N}gpu.err; echo "ref exit $?"
tail -c 600 gpurun_out/bench_ref_${N}gpu.json
