#!/bin/bash
mkdir -p gpurun_out
# 1. full GPU test-suite with timing
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/pytest_gpu_full.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_full.log
tail -14 gpurun_out/pytest_gpu_full.log
# 2. HBM copy / fill rates next to K1 (context for the roofline fraction)
python - <<'PY' > gpurun_out/hbm_rates.json 2> gpurun_out/hbm_rates.err
import torch, json
dev = torch.device("cuda", 0)
n = 12 << 30
a = torch.empty(n, dtype=torch.uint8, device=dev); b = torch.empty(n, dtype=torch.uint8, device=dev)
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
ms_copy = t(lambda: b.copy_(a)); ms_fill = t(lambda: b.zero_()); ms_read = t(lambda: a.view(torch.int64).sum())
print(json.dumps({"bytes": n, "copy_gbs_read_plus_write": 2 * n / ms_copy / 1e6, "fill_gbs_write_only": n / ms_fill / 1e6, "reduce_gbs_read_only": n / ms_read / 1e6}))
PY
cat gpurun_out/hbm_rates.json; tail -2 gpurun_out/hbm_rates.err
# 3. default bench, then ncu launch list of the same command (short variant) and the full capture of k1_bulk at C2 size
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?"
tail -c 1500 gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --octomap-scans 0"
timeout 600 $CMD > gpurun_out/plain_k1full.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k1_bulk -s 3 -c 1 -o gpurun_out/k1_prof_c2 -f $CMD > gpurun_out/ncu_k1c2.log 2>&1
tail -2 gpurun_out/ncu_k1c2.log
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --octomap-scans 8"
timeout 600 $CMD > gpurun_out/plain_l.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_default.csv $CMD > gpurun_out/ncu_ld.log 2>&1
tail -2 gpurun_out/ncu_ld.log
