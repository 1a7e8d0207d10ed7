"""Bare pinned-host <-> device copy rates of the box, the ceiling of bench.py's e2e figure (whose read-back is 12 of the
14 bytes per pixel): every rank copies 1 GB blocks device -> pinned host (and, separately, both directions at once) at
the same time; the aggregate over the ranks is what the node's PCIe / host memory system gives.

    python tools/pcie_probe.py                                   (one GPU)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/pcie_probe.py

Rank 0 prints one JSON object {"<N>": {...}}; the results of N = 1, 2, 4, 8 are merged into profiles/r2_pcie_probe.json,
which bench.py reads to report e2e as a fraction of the measured ceiling."""
import json
import os
import time

import torch

rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)

n = 1 << 30
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
h_in = torch.empty(n // 6, dtype=torch.uint8).pin_memory()
d_out = torch.empty(n, dtype=torch.uint8, device=dev)
d_in = torch.empty(n // 6, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def d2h():
    with torch.cuda.stream(s1):
        h_out.copy_(d_out, non_blocking=True)


def h2d():
    with torch.cuda.stream(s2):
        d_in.copy_(h_in, non_blocking=True)


def both():
    d2h()
    h2d()


def wall(fn, reps=6):
    fn()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    sec = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([sec], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sec = float(t.item())
    return sec / reps


t_d2h, t_h2d, t_both = wall(d2h), wall(h2d), wall(both)
if rank == 0:
    print(json.dumps({str(world): {"ranks": world, "block_bytes": n, "d2h_gbs_aggregate": world * n / t_d2h / 1e9, "h2d_gbs_aggregate": world * (n // 6) / t_h2d / 1e9,
                                   "d2h_gbs_aggregate_with_h2d_beside": world * n / t_both / 1e9, "d2h_gbs_per_rank": n / t_d2h / 1e9,
                                   "host_cpus": os.cpu_count(), "timing": "wall clock between barriers, max over ranks, 6 x 1 GB per rank"}}))
if dist is not None:
    dist.destroy_process_group()
