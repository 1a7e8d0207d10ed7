#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench exit $?"
cat gpurun_out/bench_full.json; tail -5 gpurun_out/bench_full.err
CMD="python bench.py --frames 64 --steps 2 --warmup 3 --no-cpu-baseline --octomap-scans 6"
timeout 600 $CMD > gpurun_out/plain3.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_scan_raycast -s 4 -c 1 -o gpurun_out/k3_prof -f $CMD > gpurun_out/ncu_k3.log 2>&1
timeout 600 $CMD > gpurun_out/plain3b.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 120 --csv --log-file gpurun_out/launches_octo.csv $CMD > gpurun_out/ncu_l3.log 2>&1
tail -3 gpurun_out/ncu_k3.log
ls -la gpurun_out | tail -6
