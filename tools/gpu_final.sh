#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/pytest_gpu_full.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_full.log
tail -10 gpurun_out/pytest_gpu_full.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_default.json'))
print('value',d['value'],'frac',d['roofline']['frac'],'traffic',d['roofline']['traffic'],'e2e',d['e2e']['value'],'launches',d['gpu_launches'],'clocks',d['clocks'])
print('cpu',d['cpu_baseline'])
o=d['octomap']; print('octo',o['value'],o['ms_per_scan'],o['parity_bt_ok'],'un',o['update_node']['value'],o['update_node']['parity_bt_ok'])
print('compact',d['compact_mode']); print('png',d['png_decode'])
r=json.load(open('gpurun_out/bench_ref.json')); print('ref',r['value'],r['cpu_baseline']['cores'])
PY
tail -3 gpurun_out/bench_default.err
