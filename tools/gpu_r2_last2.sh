#!/bin/bash
# last one-GPU records of the round: the scaling command at N = 1, then the default bench (with the tmpfs leg of script_e2e)
mkdir -p gpurun_out
timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 --frames 2048 --quick > gpurun_out/r2_bench_1gpu_frames2048.json 2> gpurun_out/r2_bench_1gpu_frames2048.err; echo "N=1 scaling command exit $?"
timeout 1200 python bench.py > gpurun_out/r2_bench_c2_default.json 2> gpurun_out/final_bench.err; echo "bench exit $?"; tail -3 gpurun_out/final_bench.err
python - <<'PY'
import json
for f in ('gpurun_out/r2_bench_1gpu_frames2048.json','gpurun_out/r2_bench_c2_default.json'):
    try:
        d=json.loads([l for l in open(f).read().splitlines() if l.startswith('{')][-1])
        o=d['octomap']
        print(f,'value',round(d['value']/1e9,1),'frac',round(d['roofline']['frac'],3),'e2e',round(d['e2e']['value']/1e9,2),'octo',round(o['value']),[round(x,3) for x in o['ms_per_scan_runs']],o['bt_sha256'][:12],'script',d.get('script_e2e') and (round(d['script_e2e']['frames_per_s'],1), d['script_e2e'].get('on_tmpfs')))
    except Exception as e:
        print(f,'parse failed',e)
PY
