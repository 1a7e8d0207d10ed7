#!/bin/bash
# The profiling recipe behind profiles/ (run under gpurun on one B200; ncu only after the same command exited 0 plainly).
#   gpurun --timeout 2400 -- bash tools/gpu_profile.sh
# Reads back here with:  python tools/ncu_summary.py gpurun_out/<name>.ncu-rep profiles/<name>.json
mkdir -p gpurun_out
# 1. K1 at the bench workload (C2, 4500 frames): one launch, full set
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --octomap-scans 0"
timeout 600 $CMD > gpurun_out/plain_k1.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k1_bulk -s 3 -c 1 -o gpurun_out/k1_prof_c2 -f $CMD > gpurun_out/ncu_k1.log 2>&1
# 2. the ray-casting kernel on one KITTI-shape scan of the bench sequence
CMD="python bench.py --frames 64 --steps 2 --warmup 3 --no-cpu-baseline --octomap-scans 6"
timeout 600 $CMD > gpurun_out/plain_k3.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_scan_raycast -s 4 -c 1 -o gpurun_out/k3_prof -f $CMD > gpurun_out/ncu_k3.log 2>&1
# 3. every launch of a short bench run with its device time (shares, not absolutes)
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --octomap-scans 8"
timeout 600 $CMD > gpurun_out/plain_l.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_default.csv $CMD > gpurun_out/ncu_l.log 2>&1
ls -la gpurun_out | tail -8
