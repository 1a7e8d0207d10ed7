#!/usr/bin/env python
"""K1 probe: device-resident C2 pass (4 500 KITTI-shape frames), kernel time from the library's own CUDA events, and a
bit-exact check of sampled frames against the oracle.  Used for kernel variants built with R3D_NVCC_EXTRA:

    R3D_NVCC_EXTRA="-DK1V_GROUPS=1" python 3d_reconstruction_system_b200/build.py --force && python tools/k1_probe.py
"""
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (synthetic data generator of the bench)


def main():
    import torch
    from oracle import points_oracle as po
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else bench.N_FRAMES_C2
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    r3d = importlib.import_module("3d_reconstruction_system_b200")
    ctx = r3d.Context(0)
    dev = torch.device("cuda:0")
    W, H = bench.W, bench.H
    depth, q, t = bench.synth_on_device(torch, frames, 0, dev)
    rt_host = ctx.pose_to_rt(q, t)
    rt = torch.from_numpy(rt_host).to(dev)
    out = torch.empty((frames * H * W, 3), dtype=torch.float32, device=dev)
    counts = np.zeros(frames, np.uint64)

    def step():
        ctx.backproject(depth, po.KITTI_INTRINSICS, rt=rt, depth_scale=bench.DEPTH_SCALE, out=out, shape=(frames, H, W), counts=counts)
        return ctx.last_kernel_ms()

    for _ in range(3):
        step()
    ms = sorted(step() for _ in range(reps))
    ok = True
    for k in sorted({0, 1, frames // 2, frames - 1}):
        d = depth[k].cpu().numpy().view(np.uint16)
        ref = po.depth_to_world(d, po.KITTI_INTRINSICS, rt_host[k, :9].reshape(3, 3), rt_host[k, 9:], 0, bench.DEPTH_SCALE)[1].astype(np.float32)
        got = out[k * H * W:(k + 1) * H * W].cpu().numpy()
        ok = ok and np.array_equal(got, ref)
    px = frames * H * W
    peak, _ = bench.load_peaks()
    med = ms[len(ms) // 2]
    print(json.dumps({"variant": os.environ.get("R3D_NVCC_EXTRA", ""), "frames": frames, "ms_median": med, "ms_min": ms[0],
                      "gpoints_per_s": px / med / 1e6, "gbs": px * 14 / med / 1e6, "frac": px * 14 / med / 1e6 / peak, "parity_ok": bool(ok)}))
    ctx.close()


if __name__ == "__main__":
    main()
