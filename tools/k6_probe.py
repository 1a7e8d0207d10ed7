"""K6 / compaction probe: the bench's secondary sections alone, on a short sequence (iteration tool, not a bench value).
usage: python tools/k6_probe.py [frames] [--compact] [--cpu]"""
import importlib
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    import torch
    n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 64
    r3d = importlib.import_module("3d_reconstruction_system_b200")
    cfg = bench.cfg_of("c2")
    cfg["depth_kind"] = "street"
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    ctx = r3d.Context(0)
    seq = bench.Sequence(torch, cfg, dev)
    frames = max(n, 1024 if "--compact" in sys.argv else n)
    depth = seq.frames(0, frames)
    q, t = seq.poses(0, frames)
    rt = torch.from_numpy(ctx.pose_to_rt(q, t)).to(dev)
    out = {"text_rows": bench.text_section(torch, ctx, dev, cfg, depth, rt, n=n, with_cpu="--cpu" in sys.argv)}
    if "--compact" in sys.argv:
        out["compact_mode"] = bench.compaction_section(torch, ctx, dev, cfg, depth, rt)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
