#!/bin/bash
# round 2, call 1: GPU test suite on the round's first commits, K3 variants on one box, PCIe probe
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/r2c1_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/r2c1_pytest.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/r2c1_pytest.log
tail -15 gpurun_out/r2c1_pytest.log
for tag in "" _v2 _rf4 _rf2v2; do
  R3D_LIB_PATH=$PWD/3d_reconstruction_system_b200/libr3d_b200$tag.so timeout 600 python bench.py --frames 64 --steps 3 --warmup 3 --no-cpu-baseline --octomap-scans 32 > gpurun_out/r2c1_k3$tag.json 2> gpurun_out/r2c1_k3$tag.err; echo "bench$tag exit $?"
  python - "$tag" <<'PY'
import json,sys
tag=sys.argv[1]
try:
    d=json.load(open('gpurun_out/r2c1_k3%s.json'%tag))['octomap']
    print('K3%-8s scans/s %6.0f  ms/scan runs %s  kernel ms %.3f  steps/s in kernel %.1f G' % (tag, d['value'], [round(x,3) for x in d['ms_per_scan_runs']], d['raycast_kernel_ms_last_scan'], d['raycast_steps_per_s_in_kernel']/1e9))
except Exception as e:
    print('K3',tag,'failed',e)
PY
done
timeout 300 python tools/pcie_probe.py > gpurun_out/r2c1_pcie.json 2> gpurun_out/r2c1_pcie.err; cat gpurun_out/r2c1_pcie.json
nproc; free -g | head -2
