#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/script_e2e_probe.py 256 > gpurun_out/script_e2e_probe.log 2>&1; tail -4 gpurun_out/script_e2e_probe.log | cut -c1-300; df -h /tmp /dev/shm | tail -3
CMD="python bench.py --frames 1024 --steps 2 --warmup 3 --no-cpu-baseline --octomap-scans 64"
timeout 600 $CMD > gpurun_out/ncu_plain.json 2> gpurun_out/ncu_plain.err || { echo "plain run failed"; tail -5 gpurun_out/ncu_plain.err; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_scan_walk -s 3 -c 1 -o gpurun_out/r2_k3_walk -f $CMD > gpurun_out/ncu_k3_walk.log 2>&1; echo "ncu k3_walk exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_default.csv $CMD > gpurun_out/ncu_launches.log 2>&1; echo "launch list exit $?"
