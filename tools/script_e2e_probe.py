"""Script-level end to end alone (bench.py's script_e2e section), files on the default temp dir and on tmpfs.
usage: python tools/script_e2e_probe.py [frames]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    cfg = bench.cfg_of("c2")
    out = {}
    for name, root in (("tmp", None), ("shm", "/dev/shm"), ("tmp_again", None)):
        r = bench.script_e2e_section(cfg, n_frames=n, root=root)
        out[name] = {k: r[k] for k in ("seconds", "frames_per_s", "points_per_s", "files_under", "ply_sha256")}
        print(name, out[name], flush=True)
    print(json.dumps(out))
