"""N-rank NCCL worker (one process per GPU; launch with torchrun) for the multi-GPU occupancy merge:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/workers/nccl_octomap_worker.py [n_scans] [scans_per_rank]

Every rank ray-casts only its own scans, the brick deltas are all-gathered over NCCL, and the result -- replicated
apply and owner-partitioned apply + brick gather -- must equal a serial single-GPU insertPointCloud run byte for byte
(.bt and every voxel's float32 log-odds)."""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
r3d = importlib.import_module("3d_reconstruction_system_b200")
octomap = importlib.import_module("3d_reconstruction_system_b200.octomap")
sharding = importlib.import_module("3d_reconstruction_system_b200.sharding")


def make_scan(s, n=20000):
    rng = np.random.default_rng(77 + s)
    origin = np.array([0.4 * s, 0.1 * np.sin(s), 0.0])
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    pts = origin + d * rng.uniform(0.5, 25.0, size=(n, 1))
    return pts.astype(np.float32), origin


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    n_scans = int(sys.argv[1]) if len(sys.argv) > 1 else 13
    per_rank = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = r3d.Context(local)
    res, maxrange = 0.1, 20.0
    scans = {}

    def get_scan(s):
        if s not in scans:
            scans[s] = make_scan(s)
        return scans[s]

    serial = octomap.OcTree(res, ctx=ctx)
    for s in range(n_scans):
        p, o = make_scan(s)
        serial.insertPointCloud(p, o, maxrange=maxrange)
    want_bt = serial.writeBinary()
    wk, wv = serial.voxels()

    def get_scan_batch(first, n):
        got = [get_scan(first + i) for i in range(n)]
        return np.concatenate([g[0] for g in got]), [g[0].shape[0] for g in got], np.stack([g[1] for g in got])

    def get_scan_batch_device(first, n):
        # device-resident scans: the path on which a round's ray casting starts beside the sorted apply of the round before
        pts, counts, origins = get_scan_batch(first, n)
        return torch.from_numpy(pts).cuda(), counts, origins

    for owner, batched in ((False, False), (True, False), (True, True), (False, True), (True, "device"), (False, "device")):
        tree = octomap.OcTree(res, ctx=ctx)
        sh = sharding.OctreeSharder(tree, get_scan, maxrange=maxrange, owner_partition=owner, rank=rank, world=world,
                                    get_scan_batch=get_scan_batch_device if batched == "device" else (get_scan_batch if batched else None))
        scans.clear()
        sh.run(n_scans, scans_per_rank=per_rank)
        mine = sorted(scans)
        expect = [s for _, parts in sharding.scan_rounds(n_scans, world, per_rank) for r, a, n in parts if r == rank for s in range(a, a + n)]
        assert mine == expect, (mine, expect)
        if owner:
            before = tree.numBricks()
            sharding.gather_bricks(tree)
            assert world == 1 or tree.numBricks() > before
        k, v = tree.voxels()
        assert np.array_equal(k, wk) and np.array_equal(v.view(np.uint32), wv.view(np.uint32)), "rank %d owner=%s batched=%s voxels differ" % (rank, owner, batched)
        assert tree.writeBinary() == want_bt, "rank %d owner=%s batched=%s .bt differs" % (rank, owner, batched)
    dist.barrier()
    dist.destroy_process_group()
    print("rank %d/%d ok: %d scans, %d voxels, bt %d bytes" % (rank, world, n_scans, wk.shape[0], len(want_bt)))


if __name__ == "__main__":
    main()
