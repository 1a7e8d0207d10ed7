"""K3 ray-order probe (iteration tool, not a bench value): the same scans with the points of every scan permuted into
pixel tiles / Morton order before insertPointClouds.  The tree must not change (a scan is a set of rays); the time per scan
shows how much the walk depends on which rays share a warp.
usage: python tools/k3_order_probe.py [scans]"""
import hashlib
import importlib
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def tile_perm(H, W, th, tw):
    v, u = np.mgrid[0:H, 0:W]
    key = ((v // th) * ((W + tw - 1) // tw) + (u // tw)) * (th * tw) + (v % th) * tw + (u % tw)
    return np.argsort(key.ravel(), kind="stable")


def morton_perm(H, W):
    v, u = np.mgrid[0:H, 0:W]

    def spread(x):
        x = x.astype(np.uint64)
        out = np.zeros_like(x)
        for b in range(12):
            out |= ((x >> np.uint64(b)) & np.uint64(1)) << np.uint64(2 * b)
        return out
    return np.argsort((spread(u.ravel()) | (spread(v.ravel()) << np.uint64(1))), kind="stable")


def main():
    import torch
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    r3d = importlib.import_module("3d_reconstruction_system_b200")
    octomap = importlib.import_module("3d_reconstruction_system_b200.octomap")
    cfg = bench.cfg_of("c2")
    cfg["depth_kind"] = "street"
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    ctx = r3d.Context(0)
    seq = bench.Sequence(torch, cfg, dev)
    H, W = cfg["H"], cfg["W"]
    q, t = seq.poses(512, S)
    rt = ctx.pose_to_rt(q, t)
    pts = torch.empty((S * H * W, 3), dtype=torch.float32, device=dev)
    bench.k1_call(ctx, cfg, seq.frames(512, S), torch.from_numpy(rt).to(dev), pts, S, np.zeros(S, np.uint64))
    org = bench.centres(rt)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=dev)
    orders = {"row": None, "tile8x4": tile_perm(H, W, 4, 8), "tile4x8": tile_perm(H, W, 8, 4), "tile16x2": tile_perm(H, W, 2, 16),
              "tile16x16": tile_perm(H, W, 16, 16), "morton": morton_perm(H, W), "colmajor": tile_perm(H, W, H, 1)}
    out = {}
    for name, perm in orders.items():
        p = pts if perm is None else pts.view(S, H * W, 3)[:, torch.from_numpy(perm).to(dev), :].contiguous().view(-1, 3)
        ms = []
        sha = None
        for _ in range(3):
            tree = octomap.OcTree(cfg["res"], ctx=ctx)
            tree.reserve(1 << 20)
            ctx.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            tree.insertPointClouds(p, org, maxrange=cfg["maxrange"])
            e1.record(stream)
            ctx.synchronize()
            ms.append(e0.elapsed_time(e1) / S)
            sha = hashlib.sha256(tree.writeBinary()).hexdigest()[:12]
            del tree
        out[name] = {"ms_per_scan": [round(m, 4) for m in ms], "bt": sha}
        print(name, out[name], flush=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
