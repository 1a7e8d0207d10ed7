"""Ad-hoc probe: r3d_backproject_rt with pinned host buffers (the bench's e2e figure) for the staging knobs
R3D_STAGE_SLOTS / R3D_STAGE_CHUNK_MB (read when the context is created)."""
import ctypes as C
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import points_oracle as po  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
r3d = importlib.import_module("3d_reconstruction_system_b200")
ctx = r3d.Context(0)
H, W = bench.H, bench.W
rng = np.random.default_rng(1)
lib = ctx.lib
h_in, h_out = lib.r3d_host_alloc(frames * H * W * 2), lib.r3d_host_alloc(frames * H * W * 12)
np_in = np.ctypeslib.as_array(C.cast(h_in, C.POINTER(C.c_uint16)), shape=(frames, H, W))
np_in[:] = rng.integers(0, 65535, size=(64, H, W)).astype(np.uint16)[np.arange(frames) % 64]
np_out = np.ctypeslib.as_array(C.cast(h_out, C.POINTER(C.c_float)), shape=(frames * H * W, 3))
poses = [po.synth_pose(k, frames) for k in range(frames)]
rt = ctx.pose_to_rt(np.stack([p[0] for p in poses]), np.stack([p[1] for p in poses]))
counts = np.zeros(frames, np.uint64)
ctx.backproject(np_in, po.KITTI_INTRINSICS, rt=rt, depth_scale=bench.DEPTH_SCALE, out=np_out, counts=counts)
ts = []
for _ in range(4):
    t0 = time.perf_counter()
    ctx.backproject(np_in, po.KITTI_INTRINSICS, rt=rt, depth_scale=bench.DEPTH_SCALE, out=np_out, counts=counts)
    ts.append(time.perf_counter() - t0)
ok = True
for k in sorted({0, 1, frames // 7, frames // 3, frames // 2, 2 * frames // 3, frames - 2, frames - 1}):   # every slot of the ring
    ref = po.depth_to_world(np_in[k], po.KITTI_INTRINSICS, rt[k, :9].reshape(3, 3), rt[k, 9:], 0, bench.DEPTH_SCALE)[1].astype(np.float32)
    ok = ok and bool(np.array_equal(np_out[k * H * W:(k + 1) * H * W], ref))
t = min(ts)
print(json.dumps({"slots": os.environ.get("R3D_STAGE_SLOTS"), "chunk_mb": os.environ.get("R3D_STAGE_CHUNK_MB"), "frames": frames,
                  "gpoints_per_s": frames * H * W / t / 1e9, "d2h_gbs": frames * H * W * 12 / t / 1e9, "parity_ok": ok}))
lib.r3d_host_free(h_in); lib.r3d_host_free(h_out)
ctx.close()
