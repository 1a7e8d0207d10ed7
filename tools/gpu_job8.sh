#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --frames 64 --steps 2 --warmup 3 --no-cpu-baseline --octomap-scans 6"
timeout 600 $CMD > gpurun_out/plain_k3v6.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_scan_raycast -s 4 -c 1 -o gpurun_out/k3_prof_v6 -f $CMD > gpurun_out/ncu_k3v6.log 2>&1
tail -2 gpurun_out/ncu_k3v6.log
