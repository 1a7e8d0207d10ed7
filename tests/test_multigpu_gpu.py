"""Multi-GPU occupancy merge on the GPU.  The partition / brick-merge arithmetic is checked on ONE GPU by playing the ranks
one after the other (no collectives, no concurrently waiting kernels); the NCCL path runs when the box has >= 2 GPUs
(`gpurun --gpus 2 -- python -m pytest tests/test_multigpu_gpu.py -m gpu`)."""
import importlib
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from oracle import octomap_oracle as oo

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def octomap(r3d):
    return importlib.import_module("3d_reconstruction_system_b200.octomap")


def _scan(s, n=6000):
    rng = np.random.default_rng(500 + s)
    origin = np.array([0.3 * s, -0.2 * s, 0.1])
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return (origin + d * rng.uniform(0.3, 12.0, size=(n, 1))).astype(np.float32), origin


@pytest.mark.parametrize("nparts", [2, 3, 8])
def test_owner_partition_and_brick_merge_equal_serial(octomap, r3d, nparts):
    sharding = importlib.import_module("3d_reconstruction_system_b200.sharding")
    n_scans, res, maxrange = 7, 0.1, 10.0
    serial = octomap.OcTree(res)
    ref = oo.OcTree(res)
    caster = octomap.OcTree(res)
    parts = [octomap.OcTree(res) for _ in range(nparts)]
    for s in range(n_scans):
        p, o = _scan(s)
        serial.insertPointCloud(p, o, maxrange=maxrange)
        ref.insertPointCloud_f32(p, o, maxrange)
        rec = caster.computeScanDelta(p, o, maxrange=maxrange)
        keys = rec.view(np.uint64).reshape(rec.shape[0], -1)[:, 0]
        own = sharding.brick_owner(keys, nparts)
        for r, t in enumerate(parts):
            t.applyDeltaOwned(rec, rec.shape[0], r, nparts)
            if s == 0:   # the kernel's ownership test is the host function
                assert t.numBricks() == np.count_nonzero(own == r)
    assert caster.numVoxels() == 0
    counts = [t.numBricks() for t in parts]
    assert sum(counts) == serial.numBricks() and min(counts) > 0
    merged = parts[0]
    for t in parts[1:]:
        merged.importBricks(t.exportBricks())
    k, v = merged.voxels()
    wk, wv = serial.voxels()
    assert np.array_equal(k, wk) and np.array_equal(v.view(np.uint32), wv.view(np.uint32))
    assert merged.size() == serial.size() == ref.size()
    assert merged.writeBinary() == serial.writeBinary() == ref.write_binary_bytes()


@pytest.mark.parametrize("nparts", [1, 3])
def test_sorted_round_apply_equals_scan_by_scan(octomap, r3d, nparts):
    """r3d_round.cu: the deltas of a whole round noted with deferDeltasOwned and applied in ONE sorted, scan-ordered pass
    (index -> radix sort by (brick, scan) -> one warp per brick) equal the scan-by-scan applies and the serial
    insertPointCloud run, for every voxel's float32 log-odds -- including bricks that saturate inside the round (the clamp
    makes the order matter), scans without records, a second round into the same trees and the flush by a map read."""
    ctx = r3d.default_context(0)
    res, maxrange = 0.1, 10.0
    n_scans = 23
    serial, ref = octomap.OcTree(res), oo.OcTree(res)
    caster = octomap.OcTree(res)
    recs = []
    for s in range(n_scans):
        p, o = _scan(s % 5, n=3000) if s % 7 != 6 else (np.zeros((0, 3), np.float32), np.zeros(3))   # repeats saturate voxels; one empty scan
        serial.insertPointCloud(p, o, maxrange=maxrange)
        ref.insertPointCloud_f32(p, o, maxrange)
        recs.append(caster.computeScanDelta(p, o, maxrange=maxrange))
    parts = [octomap.OcTree(res) for _ in range(nparts)]
    for lo, hi in ((0, 9), (9, n_scans)):                                   # two rounds
        flat = np.concatenate([r.reshape(-1) for r in recs[lo:hi]]) if any(r.size for r in recs[lo:hi]) else np.zeros(0, np.uint8)
        dev = ctx.to_device(flat if flat.size else np.zeros(136, np.uint8))
        counts = [r.shape[0] for r in recs[lo:hi]]
        for r, t in enumerate(parts):
            # as the merge does: one job per sending rank (here: the round cut in two pieces), all applied in one pass
            half = (hi - lo) // 2
            n0 = sum(counts[:half])
            t.deferDeltasOwned(dev.data_ptr(), counts[:half], r, nparts)
            t.deferDeltasOwned(dev.data_ptr() + n0 * 136, counts[half:], r, nparts)
            if lo == 0:
                t.flushDeferred()
            else:
                assert t.numBricks() > 0                                    # a map read flushes what was noted
        ctx.synchronize()
        dev.free()
    merged = parts[0]
    for t in parts[1:]:
        merged.importBricks(t.exportBricks())
    k, v = merged.voxels()
    wk, wv = serial.voxels()
    assert np.array_equal(k, wk) and np.array_equal(v.view(np.uint32), wv.view(np.uint32))
    assert merged.writeBinary() == serial.writeBinary() == ref.write_binary_bytes()


def test_brick_export_import_roundtrip_and_overwrite(octomap):
    a = octomap.OcTree(0.05)
    pts = np.random.default_rng(3).normal(scale=2.0, size=(20000, 3))
    a.updateNodes(pts, True)
    a.updateNodes(pts[:5000], False)
    rec = a.exportBricks()
    assert rec.shape == (a.numBricks(), 2120)
    b = octomap.OcTree(0.05)
    b.updateNodes(pts[:100], True)          # overlapping bricks already present: known voxels are replaced
    b.importBricks(rec)
    ka, va = a.voxels()
    kb, vb = b.voxels()
    assert np.array_equal(ka, kb) and np.array_equal(va.view(np.uint32), vb.view(np.uint32))
    assert a.writeBinary() == b.writeBinary()
    c = octomap.OcTree(0.05)
    c.importBricks(rec[:0])
    assert c.numVoxels() == 0


def test_nccl_merge_two_ranks():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (run under gpurun --gpus 2)")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "workers", "nccl_octomap_worker.py"), "13", "2"]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-4000:] + p.stderr[-4000:]
    assert "rank 0/2 ok" in p.stdout and "rank 1/2 ok" in p.stdout
