"""Ad-hoc probe: throughput of K1's compaction mode (device resident) next to the plain mode."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
r3d = importlib.import_module("3d_reconstruction_system_b200")
import bench
from oracle import points_oracle as po
ctx = r3d.Context(0)
dev = torch.device("cuda", 0)
n = 1024
depth, q, t = bench.synth_on_device(torch, n, 0, dev)
rt = torch.from_numpy(ctx.pose_to_rt(q, t)).to(dev)
out = torch.empty((n * bench.H * bench.W, 3), dtype=torch.float32, device=dev)
cnt = torch.zeros(n, dtype=torch.int64, device=dev)
stream = torch.cuda.ExternalStream(ctx.stream(), device=dev)
for compact in (False, True):
    for _ in range(3):
        ctx.backproject(depth, po.KITTI_INTRINSICS, rt=rt, depth_scale=1 / 256.0, out=out, shape=(n, bench.H, bench.W), counts=cnt, compact=compact)
    ctx.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(5):
        ctx.backproject(depth, po.KITTI_INTRINSICS, rt=rt, depth_scale=1 / 256.0, out=out, shape=(n, bench.H, bench.W), counts=cnt, compact=compact)
    e1.record(stream)
    ctx.synchronize()
    ms = e0.elapsed_time(e1) / 5
    px = n * bench.H * bench.W
    valid = int(cnt.sum().item())
    print("compact=%s: %.3f ms, %.1f G input px/s, %d of %d valid (%.1f%%), %.0f GB/s algorithmic" % (
        compact, ms, px / ms / 1e6, valid, px, 100.0 * valid / px, (px * 2 + valid * 12) / ms / 1e6))
