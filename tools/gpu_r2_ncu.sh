#!/bin/bash
# ncu ledger of round 2: one --set full capture per kernel family at bench-like sizes (run only after the same command has
# exited 0 without ncu), plus the per-launch time list of a short default run.  Read back with tools/ncu_summary.py.
mkdir -p gpurun_out
timeout 900 python bench.py --config c3 --steps 3 > gpurun_out/r2_bench_c3.json 2> gpurun_out/r2_bench_c3.err; python -c "import json; d=json.load(open('gpurun_out/r2_bench_c3.json')); o=d['octomap']; print('c3', round(o['value']), round(o['ms_per_scan'],3), o['growth'], o['host_pipeline'], o['bt_sha256'][:12])"
CMD="python bench.py --frames 1024 --steps 2 --warmup 3 --no-cpu-baseline --octomap-scans 32"
timeout 600 $CMD > gpurun_out/ncu_plain.json 2> gpurun_out/ncu_plain.err || { echo "plain run failed"; tail -5 gpurun_out/ncu_plain.err; exit 1; }
cap() { # name regex skip count
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -o gpurun_out/r2_$1 -f $CMD > gpurun_out/ncu_$1.log 2>&1; echo "ncu $1 exit $?"
}
cap k2_update "k_points_(ensure|update)" 2 2
cap k3_prepare "k_scan_prepare" 3 1
cap k3_walk "k_scan_walk" 3 1
cap k3_list_emit "k_cells_(list|emit)" 6 2
cap k4_apply "k_apply_delta" 8 2
cap k5_bt "k_bt_" 0 12
cap k6_rows "k6_rows" 1 2
cap k1_compact "k1_(bulk_compact|count_tiles)" 2 2
cap k1_vec "k1_bulk_vec" 3 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_default.csv $CMD > gpurun_out/ncu_launches.log 2>&1; echo "launch list exit $?"
ls -la gpurun_out/*.ncu-rep | tail -12
