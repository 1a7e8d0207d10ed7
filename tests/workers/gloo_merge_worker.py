"""World-size-N gloo worker for tests/test_sharding_cpu.py (CPU only).

Drives 3d_reconstruction_system_b200.sharding.merged_insert -- the host logic of the multi-GPU occupancy merge --
with the ORACLE standing in for the GPU map (this is a test: the oracle is the checker and here also the stand-in backend;
the product's GPU backend is exercised by tests/test_multigpu_gpu.py).  Each rank ray-casts only its own scans
(oracle computeUpdate), ships them as 136-byte records over gloo, applies all records in global scan order, and checks
that its final tree equals a serial insertPointCloud run bit for bit."""
import hashlib
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sharding = importlib.import_module("3d_reconstruction_system_b200.sharding")
from oracle import octomap_oracle as oo  # noqa: E402

REC = sharding.RECORD_BYTES          # 17 uint64: [count, 16 x (key48 | occupied << 63)]
KEYS_PER_REC = 16


def make_scan(s):
    rng = np.random.default_rng(1000 + s)
    if s % 5 == 3:
        return np.zeros((0, 3), np.float32), np.array([0.05 * s, 0.0, 0.0])      # an empty scan
    n = int(rng.integers(5, 60))
    origin = np.array([0.05 * s, 0.01 * s, 0.0])
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    pts = origin + d * rng.uniform(0.2, 3.0, size=(n, 1))
    return pts.astype(np.float32), origin


def encode(free, occ):
    keys = np.concatenate([free, occ | np.uint64(1 << 63)])
    n_rec = (keys.size + KEYS_PER_REC - 1) // KEYS_PER_REC
    out = np.zeros((n_rec, 17), np.uint64)
    for r in range(n_rec):
        part = keys[r * KEYS_PER_REC:(r + 1) * KEYS_PER_REC]
        out[r, 0] = part.size
        out[r, 1:1 + part.size] = part
    return out.view(np.uint8).reshape(-1)


def decode(rec_bytes):
    a = np.frombuffer(bytes(rec_bytes), dtype=np.uint64).reshape(-1, 17)
    keys = np.concatenate([row[1:1 + int(row[0])] for row in a]) if a.shape[0] else np.zeros(0, np.uint64)
    occ = (keys >> np.uint64(63)).astype(bool)
    return keys[~occ], keys[occ] & np.uint64((1 << 63) - 1)


def apply_keys(tree, free, occ):
    # insertPointCloud's tail: every free key updateNode(key, false), then every occupied key updateNode(key, true)
    for packed, flag in ((free, False), (occ, True)):
        k = oo.unpack_keys(packed)
        for kk in k:
            tree.updateNode([tree.keyToCoord(kk[0]), tree.keyToCoord(kk[1]), tree.keyToCoord(kk[2])], flag)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    n_scans, per_rank = int(sys.argv[1]), int(sys.argv[2])
    overlap = len(sys.argv) > 3 and sys.argv[3] in ("overlap", "regrow")
    # "regrow": exchange slots that start far too small, so that the header's "did not fit" path runs (several times)
    state = {"slot_min": 136} if (len(sys.argv) > 3 and sys.argv[3] == "regrow") else {}
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % os.environ["MASTER_PORT"], rank=rank, world_size=world)
    res, maxrange = 0.1, 2.5
    caster = oo.OcTree(res)          # never updated: only computeUpdate
    tree = oo.OcTree(res)
    log = []
    cast_here = []

    def compute_delta(s, out, off):
        pts, origin = make_scan(s)
        cast_here.append(s)
        free, occ = caster.computeUpdate(pts, origin, maxrange)
        b = encode(free, occ)
        need = off + b.size
        if need > out.numel():
            grown = torch.zeros(max(need, 2 * out.numel()), dtype=torch.uint8)
            grown[:off] = out[:off]
            out = grown
        out[off:need] = torch.from_numpy(b.copy())
        return b.size // REC, out

    def apply_delta(rec, n, s):
        assert rec.numel() == n * REC
        raw = rec.numpy().tobytes()
        log.append((s, n, hashlib.sha256(raw).hexdigest()))
        free, occ = decode(raw)
        apply_keys(tree, free, occ)

    applied = sharding.merged_insert(n_scans, rank, world, compute_delta, apply_delta, lambda nb: torch.zeros(64, dtype=torch.uint8),
                                     scans_per_rank=per_rank, overlap=overlap, state=state)
    if "slot_min" in state:
        assert state.get("slot_hint", 0) > 136, "the regrow path did not run"
    # serial reference on every rank
    ref = oo.OcTree(res)
    ref_log = []
    for s in range(n_scans):
        pts, origin = make_scan(s)
        free, occ = caster.computeUpdate(pts, origin, maxrange)
        b = encode(free, occ)
        if b.size:
            ref_log.append((s, b.size // REC, hashlib.sha256(b.tobytes()).hexdigest()))
        ref.insertPointCloud_f32(pts, origin, maxrange)
    assert log == ref_log, "rank %d applied a different record sequence" % rank
    assert applied == sum(n for _, n, _ in ref_log)
    # this rank ray-cast exactly its share
    expect = [s for _, parts in sharding.scan_rounds(n_scans, world, per_rank) for r, a, n in parts if r == rank for s in range(a, a + n)]
    assert cast_here == expect
    k1, v1, d1 = tree.leaves()
    k2, v2, d2 = ref.leaves()
    assert np.array_equal(k1, k2) and np.array_equal(v1.view(np.uint32), v2.view(np.uint32)) and np.array_equal(d1, d2)
    assert tree.write_binary_bytes() == ref.write_binary_bytes()
    # frame sharding of the point path: ranges tile [0, n) in rank order
    lo, hi = sharding.frame_range(n_scans, world, rank)
    t = torch.tensor([lo, hi], dtype=torch.int64)
    allr = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(allr, t)
    edges = [int(v) for pair in allr for v in pair]
    assert edges[0] == 0 and edges[-1] == n_scans and all(edges[2 * i + 1] == edges[2 * i + 2] for i in range(world - 1))
    dist.barrier()
    dist.destroy_process_group()
    print("rank %d ok: %d scans, %d records" % (rank, n_scans, applied))


if __name__ == "__main__":
    main()
