#!/bin/bash
# Re-capture of the kernel families that changed after the round's first ledger (K3 walk: two slots, 256-ray chunks; K6:
# fma units, three rows per thread; K1 compaction: all-valid tile loop), same command as tools/gpu_r2_ncu.sh.
mkdir -p gpurun_out
CMD="python bench.py --frames 1024 --steps 2 --warmup 3 --no-cpu-baseline --octomap-scans 32"
timeout 600 $CMD > gpurun_out/ncu_plain.json 2> gpurun_out/ncu_plain.err || { echo "plain run failed"; tail -5 gpurun_out/ncu_plain.err; exit 1; }
cap() { # name regex skip count
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -o gpurun_out/r2_$1 -f $CMD > gpurun_out/ncu_$1.log 2>&1; echo "ncu $1 exit $?"
}
cap k3_walk "k_scan_walk" 3 1
cap k3_prepare "k_scan_prepare" 3 1
cap k6_rows "k6_rows" 1 2
cap k1_compact "k1_(bulk_compact|count_tiles)" 2 2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_default.csv $CMD > gpurun_out/ncu_launches.log 2>&1; echo "launch list exit $?"
ls -la gpurun_out/r2_k3_walk.ncu-rep gpurun_out/r2_k6_rows.ncu-rep gpurun_out/r2_k1_compact.ncu-rep
