"""Parity of the GPU occupancy path (updateNode / insertPointCloud / writeBinary) with the OctoMap restatement
in oracle/octomap_oracle.c, through the drop-in octomap.OcTree class (which calls the C ABI).

Bars: voxel key sets and occupancy bits bit-exact, log-odds bit-exact (stricter than the 1e-6 of BASELINE.json),
.bt byte-identical.  PARITY UNPINNED vs real OctoMap (not installable here): the oracle is the spec."""
import importlib
import struct

import numpy as np
import pytest

from oracle import octomap_oracle as oo
from oracle import points_oracle as po

pytestmark = pytest.mark.gpu

HDR = (b"# Octomap OcTree binary file\n# (feel free to add / change comments, but leave the first line as it is!)\n#\n"
       b"id OcTree\nsize %d\nres %s\ndata\n")


@pytest.fixture(scope="module")
def octomap(r3d):
    return importlib.import_module("3d_reconstruction_system_b200.octomap")


def f32bits(v):
    return struct.unpack("<I", struct.pack("<f", float(v)))[0]


def expand_leaves(rk, rv, rd):
    """The oracle's (possibly pruned) leaves -> every depth-16 voxel they stand for: (packed keys sorted, values)."""
    keys, vals = [], []
    for d in np.unique(rd):
        m = rd == d
        side = 1 << (16 - int(d))
        base = rk[m].astype(np.int64)
        off = np.arange(side, dtype=np.int64)
        ox, oy, oz = np.meshgrid(off, off, off, indexing="ij")
        offs = np.stack([ox.ravel(), oy.ravel(), oz.ravel()], axis=1)
        k = (base[:, None, :] + offs[None, :, :]).reshape(-1, 3).astype(np.uint64)
        keys.append(k[:, 0] | (k[:, 1] << np.uint64(16)) | (k[:, 2] << np.uint64(32)))
        vals.append(np.repeat(rv[m], side ** 3))
    if not keys:
        return np.zeros(0, np.uint64), np.zeros(0, np.float32)
    keys, vals = np.concatenate(keys), np.concatenate(vals)
    order = np.argsort(keys, kind="stable")
    return keys[order], vals[order]


def assert_same_tree(gpu, ref, check_size=True):
    """EVERY voxel of the GPU map has the oracle's key and log-odds bit for bit (the whole arrays are compared, not a
    sample), the voxel counts agree, and the .bt bytes agree."""
    keys, vals = gpu.voxels()                      # sorted by packed key
    rkeys, rvals = expand_leaves(*ref.leaves())
    packed = oo.pack_keys(keys) if keys.shape[0] else np.zeros(0, np.uint64)
    assert packed.shape[0] == rkeys.shape[0]
    assert np.array_equal(packed, rkeys)
    bad = np.flatnonzero(vals.view(np.uint32) != rvals.view(np.uint32))
    assert bad.size == 0, (bad.size, keys[bad[:3]], vals[bad[:3]], rvals[bad[:3]])
    if check_size:
        assert gpu.size() == ref.size()
    assert gpu.writeBinary() == ref.write_binary_bytes()   # oracle call last: it mutates the oracle tree


def test_params_and_keys(octomap):
    t = octomap.OcTree(0.1)
    r = oo.OcTree(0.1)
    assert [f32bits(v) for v in (t._hit, t._miss, t._cmin, t._cmax)] == [0x3F58E883, 0xBECF991F, 0xC0000075, 0x4060B4BA]
    rng = np.random.default_rng(0)
    pts = np.concatenate([rng.uniform(-3300, 3300, size=(20000, 3)),
                          np.array([[0, 0, 0], [-1e-9, 0.05, 0.1], [0.3, 0.7, -0.3], [3276.75, 0, 0], [3276.8, 0, 0], [-3276.8, 0, 0],
                                    [np.nan, 0, 0], [np.inf, 1, 1], [1e30, 0, 0]])])
    k = np.arange(-40, 40)[:, None] * 0.1 + np.array([0.0, 1e-7, -1e-7])[None, :]
    pts = np.concatenate([pts, np.stack([k.ravel(), k.ravel(), k.ravel()], axis=1)])
    keys, valid = t.coordsToKeys(pts)
    for p, kk, v in zip(pts, keys, valid):
        want = r.coordToKey(np.float32(p).astype(np.float64))
        assert (want is not None) == bool(v)
        if want is not None:
            assert tuple(int(x) for x in kk) == want
    t5 = octomap.OcTree(0.05)
    assert t5.coordToKey([1638.39, 0, 0])[0] == 65535 and t5.coordToKey([1638.41, 0, 0]) is None


def test_logodds_ladder_and_float_update(octomap):
    t, r = octomap.OcTree(0.1), oo.OcTree(0.1)
    p = np.array([1.0, 2.0, 3.0])
    k = r.coordToKey(p)
    for i in range(7):
        t.updateNode(p, True)
        r.updateNode(p, True)
        assert f32bits(t.search(k)) == f32bits(r.search(k))
    for i in range(20):
        t.updateNode(p, False)
        r.updateNode(p, False)
        if i % 3 == 0:
            assert f32bits(t.search(k)) == f32bits(r.search(k))
    assert f32bits(t.search(k)) == 0xC0000075
    t.updateNode(p, 1.0)
    r.updateNode(p, 1.0)
    t.updateNode(p, -0.25)
    r.updateNode(p, -0.25)
    assert f32bits(t.search(k)) == f32bits(r.search(k))
    assert t.search((1, 2, 3)) is None
    # many hits on one voxel in ONE batch (warp-aggregated path)
    t2, r2 = octomap.OcTree(0.1), oo.OcTree(0.1)
    for n in (1, 2, 3, 4, 5, 33, 1000):
        t2.clear()
        r2 = oo.OcTree(0.1)
        pts = np.tile(p, (n, 1))
        t2.updateNodes(pts, True)
        r2.updateNodes(pts, True)
        assert f32bits(t2.search(k)) == f32bits(r2.search(k))


def test_bt_known_answers(octomap):
    t = octomap.OcTree(0.1)
    assert t.writeBinary() == HDR % (0, b"0.1")
    assert t.size() == 0
    t.updateNode(np.array([0.05, 0.05, 0.05]), True)
    t.updateInnerOccupancy()
    assert t.size() == 17
    assert t.writeBinary() == HDR % (17, b"0.1") + b"\x00\xc0" + b"\x03\x00" * 14 + b"\x02\x00"
    t = octomap.OcTree(0.05)
    t.updateNode(np.array([-0.01, -0.01, -0.01]), False)
    assert t.writeBinary() == HDR % (17, b"0.05") + b"\x03\x00" + b"\x00\xc0" * 14 + b"\x00\x40"
    t = octomap.OcTree(0.1)
    for dx in (0, 1):
        for dy in (0, 1):
            for dz in (0, 1):
                t.updateNode(np.array([0.05 + 0.1 * dx, 0.05 + 0.1 * dy, 0.05 + 0.1 * dz]), True)
    assert t.size() == 16
    assert t.writeBinary() == HDR % (16, b"0.1") + b"\x00\xc0" + b"\x03\x00" * 13 + b"\x02\x00"
    t = octomap.OcTree(0.1)
    t.updateNode(np.array([0.05, 0.05, 0.05]), True)
    t.updateNode(np.array([0.15, 0.05, 0.05]), False)
    t.updateNode(np.array([0.05, 0.15, 0.15]), True)
    assert t.writeBinary() == HDR % (19, b"0.1") + b"\x00\xc0" + b"\x03\x00" * 14 + bytes([0x02 | 0x04, 0x20])


def test_prune_early_break_quirk(octomap):
    t, t2 = octomap.OcTree(0.1), octomap.OcTree(0.1)
    for ix in range(4):
        for iy in range(4):
            for iz in range(4):
                p = np.array([0.05 + 0.1 * ix, 0.05 + 0.1 * iy, 0.05 + 0.1 * iz])
                t.updateNode(p, True)
                t2.updateNode(p, True)
                if not (ix < 2 and iy < 2 and iz < 2):
                    t.updateNode(p, True)
                if ix % 2 == 0:
                    t2.updateNode(p, True)
    assert t.size() == 1 + 14 + 8
    assert t.writeBinary() == HDR % (23, b"0.1") + b"\x00\xc0" + b"\x03\x00" * 13 + b"\xaa\xaa"
    assert t2.writeBinary() == HDR % (15, b"0.1") + b"\x00\xc0" + b"\x03\x00" * 12 + b"\x02\x00"


@pytest.mark.parametrize("res,n,spread,seed", [(0.1, 3000, 2.0, 1), (0.1, 60000, 1.2, 2), (0.05, 40000, 30.0, 3), (0.1, 20000, 3000.0, 4)])
def test_update_node_clouds_match_oracle(octomap, res, n, spread, seed):
    """The reference scripts' mode: updateNode(p, True) per point, then writeBinary."""
    rng = np.random.default_rng(seed)
    pts = rng.uniform(-spread, spread, size=(n, 3))
    pts[: n // 10] = np.round(pts[: n // 10], 1)     # duplicates / voxel-boundary values
    t, r = octomap.OcTree(res), oo.OcTree(res)
    t.updateNodes(pts, True)
    r.updateNodes(pts, True)
    assert_same_tree(t, r)


def test_dense_block_prunes_across_levels(octomap):
    """A solid 32^3 block (4 bricks per axis... all saturated) collapses many levels; plus a shell of mixed values."""
    g = (np.arange(32) + 0.5) * 0.1
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    pts = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    t, r = octomap.OcTree(0.1), oo.OcTree(0.1)
    for _ in range(6):   # saturate
        t.updateNodes(pts, True)
        r.updateNodes(pts, True)
    t.updateNodes(pts[::7], False)
    r.updateNodes(pts[::7], False)
    assert_same_tree(t, r)
    t2, r2 = octomap.OcTree(0.1), oo.OcTree(0.1)
    for _ in range(6):
        t2.updateNodes(pts, True)
        r2.updateNodes(pts, True)
    assert_same_tree(t2, r2)


def test_per_point_api_mixed_sequence(octomap):
    rng = np.random.default_rng(11)
    pts = np.round(rng.uniform(-0.6, 0.6, size=(4000, 3)), 2)
    occ = rng.random(4000) < 0.6
    t, r = octomap.OcTree(0.1), oo.OcTree(0.1)
    for p, o in zip(pts, occ):
        t.updateNode(p, bool(o))
        r.updateNode(p, bool(o))
    assert_same_tree(t, r)


def _scan(rng, n, origin, far=30.0):
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return origin + d * rng.uniform(0.0, far, size=(n, 1))


@pytest.mark.parametrize("maxrange", [-1.0, 12.0])
@pytest.mark.parametrize("res", [0.1, 0.05])
def test_scan_delta_keys_bit_exact(octomap, maxrange, res):
    rng = np.random.default_rng(21)
    origin = np.array([0.31, -1.27, 0.55])
    pts = _scan(rng, 3000, origin)
    pts[:5] = origin                                     # zero-length rays
    pts[5] = origin + [3.0, 0, 0]                         # axis aligned
    pts[6] = origin + [0, -2.0, 0]
    pts[7] = origin + [1e-4, 1e-4, 1e-4]
    pts[8] = [5000.0, 0, 0]                               # out of bounds endpoint
    t, r = octomap.OcTree(res), oo.OcTree(res)
    rec = t.computeScanDelta(pts, origin, maxrange)
    fk, ok = octomap.OcTree.deltaKeys(rec)
    rf, ro = r.computeUpdate(pts.astype(np.float32), origin, maxrange)
    assert np.array_equal(np.sort(oo.pack_keys(fk)), rf)
    assert np.array_equal(np.sort(oo.pack_keys(ok)), ro)
    assert t.numVoxels() == 0                             # computing a delta does not touch the tree


def test_insert_point_cloud_sequence_matches_oracle(octomap):
    rng = np.random.default_rng(22)
    t, r = octomap.OcTree(0.1), oo.OcTree(0.1)
    for s in range(8):
        origin = np.array([0.4 * s, 0.1 * s, -0.05 * s])
        pts = _scan(rng, 1500, origin, far=10.0)
        mr = -1.0 if s % 2 else 6.0
        t.insertPointCloud(pts, origin, maxrange=mr)
        r.insertPointCloud(pts, origin, maxrange=mr)
    assert_same_tree(t, r)


def test_insert_point_cloud_discretize_and_oob_origin(octomap):
    rng = np.random.default_rng(23)
    origin = np.array([1.0, 1.0, 1.0])
    pts = np.round(_scan(rng, 2000, origin, far=4.0), 1)
    t, r = octomap.OcTree(0.1), oo.OcTree(0.1)
    t.insertPointCloud(pts, origin, maxrange=-1.0, discretize=True)
    r.insertPointCloud(pts, origin, maxrange=-1.0, discretize=True)
    # sensor outside the map: no free cells, endpoints still occupied
    t.insertPointCloud(pts[:100], np.array([9000.0, 0, 0]), maxrange=-1.0)
    r.insertPointCloud(pts[:100], np.array([9000.0, 0, 0]), maxrange=-1.0)
    t.insertPointCloud(np.zeros((0, 3)), origin)
    assert_same_tree(t, r)


def test_delta_export_apply_equals_insert(octomap):
    rng = np.random.default_rng(24)
    a, b, r = octomap.OcTree(0.1), octomap.OcTree(0.1), oo.OcTree(0.1)
    for s in range(3):
        origin = np.array([0.2 * s, 0.0, 0.0])
        pts = _scan(rng, 2000, origin, far=8.0)
        a.insertPointCloud(pts, origin, maxrange=5.0)
        rec = b.computeScanDelta(pts, origin, maxrange=5.0)
        b.applyDelta(rec)
        r.insertPointCloud(pts, origin, maxrange=5.0)
    assert a.writeBinary() == b.writeBinary()
    ka, va = a.voxels()
    kb, vb = b.voxels()
    assert np.array_equal(ka, kb) and np.array_equal(va.view(np.uint32), vb.view(np.uint32))
    assert_same_tree(b, r)


def test_kitti_shape_scan_from_fused_points(octomap, r3d):
    """Config-3 shape: one 1242x375 street frame -> K1 world points (float32) -> insertPointCloud @0.1 m / 80 m."""
    ctx = r3d.default_context(0)
    intr = po.KITTI_INTRINSICS
    depth = po.synth_depth_u16(1242, 375, intr, 20261018 + 3, "street")
    q, tr = po.synth_pose(2250, 4500)
    rt = ctx.pose_to_rt(q, tr)
    world, _ = ctx.backproject(depth, intr, rt=rt, depth_scale=1 / 256.0)
    world = world[::3]                                   # keep the oracle run short (~1.5 s)
    origin = po.camera_centre(rt[0, :9].reshape(3, 3), rt[0, 9:])
    t, r = octomap.OcTree(0.1), oo.OcTree(0.1)
    rec = t.computeScanDelta(world, origin, 80.0)
    fk, ok = octomap.OcTree.deltaKeys(rec)
    rf, ro = r.computeUpdate(world, origin, 80.0)
    assert np.array_equal(np.sort(oo.pack_keys(fk)), rf)
    assert np.array_equal(np.sort(oo.pack_keys(ok)), ro)
    t.insertPointCloud(world, origin, maxrange=80.0)
    r.insertPointCloud_f32(world, origin, 80.0)
    assert_same_tree(t, r)


def test_scratch_growth_on_wide_scan(octomap):
    """A scan whose free space covers far more bricks than the initial scratch table: the table grows and re-casts."""
    rng = np.random.default_rng(25)
    origin = np.zeros(3)
    pts = _scan(rng, 60000, origin, far=120.0)
    t, r = octomap.OcTree(0.1), oo.OcTree(0.1)
    rec = t.computeScanDelta(pts, origin, -1.0)
    fk, ok = octomap.OcTree.deltaKeys(rec)
    rf, ro = r.computeUpdate(pts.astype(np.float32), origin, -1.0)
    assert np.array_equal(np.sort(oo.pack_keys(fk)), rf)
    assert np.array_equal(np.sort(oo.pack_keys(ok)), ro)


def test_dropped_points_and_device_buffers(octomap):
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(26)
    pts = rng.uniform(-5, 5, size=(5000, 3)).astype(np.float32)
    pts[:17, 0] = 4000.0
    t, r = octomap.OcTree(0.1), oo.OcTree(0.1)
    d = torch.from_numpy(pts).cuda()
    t.updateNodes(d, True)
    assert t.n_dropped == 17
    r.updateNodes_f32(pts, True)
    origin = np.array([0.0, 0.0, 0.0])
    t.insertPointCloud(d[100:2000], origin, maxrange=4.0)
    r.insertPointCloud_f32(pts[100:2000], origin, 4.0)
    assert_same_tree(t, r)


def test_read_binary_roundtrip(octomap, tmp_path):
    """writeBinary -> readBinary -> writeBinary is the identity on the bytes; every voxel comes back at the clamping value
    of its occupancy; pruned leaves (whole bricks and larger) expand to voxels."""
    rng = np.random.default_rng(31)
    t = octomap.OcTree(0.1)
    origin = np.array([0.2, -0.1, 0.3])
    t.insertPointCloud(_scan(rng, 20000, origin, far=25.0), origin, maxrange=20.0)
    # a solid block so that pruning reaches above brick level (depth 12 = 16^3 voxels)
    g = (np.arange(32) + 0.5) * 0.1 + 40.0
    t.updateNodes(np.stack(np.meshgrid(g, g, g, indexing="ij"), axis=-1).reshape(-1, 3), True)
    bt = t.writeBinary()
    p = tmp_path / "a.bt"
    assert t.writeBinary(bytes(str(p), "utf-8"))
    u = octomap.OcTree(0.5)                      # resolution comes from the file
    assert u.readBinary(bytes(str(p), "utf-8"))
    assert u.getResolution() == 0.1
    assert u.writeBinary() == bt
    k0, v0 = t.voxels()
    k1, v1 = u.voxels()
    assert np.array_equal(k0, k1)
    want = np.where(v0 >= 0, np.float32(t._cmax), np.float32(t._cmin))
    assert np.array_equal(v1.view(np.uint32), want.view(np.uint32))
    assert u.size() == oo_size_after_ml(t, octomap)
    # from memory, then keep mapping into the loaded tree
    w = octomap.OcTree(0.1)
    assert w.readBinary(bt)
    assert w.writeBinary() == bt
    w.insertPointCloud(_scan(rng, 500, origin, far=5.0), origin, maxrange=-1.0)
    assert w.numVoxels() >= u.numVoxels()
    # the oracle's files load too
    r = oo.OcTree(0.05)
    pts = rng.normal(scale=1.5, size=(5000, 3))
    r.updateNodes(pts, True)
    rb = r.write_binary_bytes()
    x = octomap.OcTree(0.1)
    assert x.readBinary(rb) and x.getResolution() == 0.05 and x.writeBinary() == rb
    assert not octomap.OcTree(0.1).readBinary(bytes(str(tmp_path / "missing.bt"), "utf-8"))
    with pytest.raises(Exception):
        octomap.OcTree(0.1).readBinary(bt[:len(bt) // 2])
    empty = octomap.OcTree(0.25).writeBinary()
    e = octomap.OcTree(0.1)
    assert e.readBinary(empty) and e.numVoxels() == 0 and e.writeBinary() == empty


def oo_size_after_ml(tree, octomap):
    """size() of the max-likelihood, pruned tree = the `size` field writeBinary puts in the header."""
    hdr = tree.writeBinary().split(b"data\n")[0].decode()
    return int([ln for ln in hdr.splitlines() if ln.startswith("size ")][0].split()[1])


def test_insert_point_clouds_batch_equals_loop(octomap):
    rng = np.random.default_rng(41)
    S, N = 5, 3000
    origins = np.array([[0.3 * s, -0.1 * s, 0.05] for s in range(S)])
    scans = np.stack([_scan(rng, N, origins[s], far=15.0) for s in range(S)]).astype(np.float32)
    a, b, r = octomap.OcTree(0.1), octomap.OcTree(0.1), oo.OcTree(0.1)
    for s in range(S):
        a.insertPointCloud(scans[s], origins[s], maxrange=12.0)
        r.insertPointCloud_f32(scans[s], origins[s], 12.0)
    b.insertPointClouds(scans, origins, maxrange=12.0)
    assert b.lastScanStats()["rays"] == S * N
    ragged = octomap.OcTree(0.1)
    ragged.insertPointClouds(np.concatenate([scans[0], scans[1][:100], scans[2]]), origins[:3], maxrange=12.0, counts=[N, 100, N])
    assert ragged.numVoxels() > 0
    assert a.writeBinary() == b.writeBinary()
    assert_same_tree(b, r)


def test_c4_airsim_disparity_octree_005(octomap, r3d):
    """BASELINE config 4: 640x480 disparity (PSMNet-style uint16/256) -> Z = fx*B/d -> world -> insertPointCloud at 0.05 m,
    80 m (the dense scan scratch is 405^3 brick cells here), against the oracle on a sub-sampled frame."""
    ctx = r3d.default_context(0)
    intr = po.AIRSIM_INTRINSICS
    W, H, B = 640, 480, 0.25
    z = po.synth_depth_u16(W, H, intr, 20261018 + 4, "street").astype(np.float64) / 256.0
    disp = np.where(z > 0, np.round(256.0 * intr[0] * B / np.maximum(z, 1e-9)), 0).clip(0, 65535).astype(np.uint16)
    q, tr = po.synth_pose(500, 1000, step=0.2)
    rt = ctx.pose_to_rt(q, tr)
    world, _ = ctx.backproject(disp, intr, rt=rt, mode=po.MODE_DISPARITY, depth_scale=1.0 / 256.0, fB=intr[0] * B)
    _, wref = po.depth_to_world(disp, intr, rt[0, :9].reshape(3, 3), rt[0, 9:], po.MODE_DISPARITY, 1.0 / 256.0, intr[0] * B)
    assert np.array_equal(world, wref.astype(np.float32))
    world = world[::5]
    origin = po.camera_centre(rt[0, :9].reshape(3, 3), rt[0, 9:])
    t, r = octomap.OcTree(0.05), oo.OcTree(0.05)
    t.insertPointCloud(world, origin, maxrange=80.0)
    r.insertPointCloud_f32(world, origin, 80.0)
    assert_same_tree(t, r)
    # idempotence of the read-only passes
    bt = t.writeBinary()
    t.updateInnerOccupancy()
    assert t.writeBinary() == bt
    t.toMaxLikelihood()
    assert t.writeBinary() == bt
    t.toMaxLikelihood()
    assert t.writeBinary() == bt


def test_insert_point_clouds_pipelined_device_batch(octomap, r3d):
    """The two-deep pipeline of r3d_tree_insert_scans (device-resident scans, bounded range): equal to the scan-by-scan
    loop and to the oracle, including a batch whose records outgrow the record buffer mid-way (abort / re-list path),
    odd and even batch sizes, empty scans and a second batch into the same tree."""
    ctx = r3d.default_context(0)
    rng = np.random.default_rng(43)
    S, N = 7, 4000
    origins = np.array([[0.25 * s, 0.1 * s, 0.0] for s in range(S)])
    scans = np.stack([_scan(rng, N, origins[s], far=(3.0 if s < 3 else 40.0)) for s in range(S)]).astype(np.float32)
    scans[4] = origins[4]                                 # a scan of zero-length rays
    dev = ctx.to_device(scans.reshape(-1, 3))
    a, b, r = octomap.OcTree(0.1), octomap.OcTree(0.1), oo.OcTree(0.1)
    for s in range(S):
        a.insertPointCloud(scans[s], origins[s], maxrange=30.0)
        r.insertPointCloud_f32(scans[s], origins[s], 30.0)
    b.insertPointClouds(dev, origins, maxrange=30.0)      # small scans first: the far ones overflow the 65 536-record start buffer? (no: forces growth only if needed)
    assert b.lastScanStats()["rays"] == S * N
    assert a.writeBinary() == b.writeBinary()
    assert_same_tree(b, r)
    # even-sized second batch into the same tree, then per-scan delta export still refers to the last scan
    c = octomap.OcTree(0.1)
    c.insertPointClouds(dev[: 4 * N], origins[:4], maxrange=30.0)
    c.insertPointClouds(dev[4 * N:6 * N], origins[4:6], maxrange=30.0)
    c.insertPointCloud(scans[6], origins[6], maxrange=30.0)
    assert c.writeBinary() == a.writeBinary()


def test_ot_write_read_match_oracle(octomap, tmp_path):
    """tree.write() (.ot: every node's log-odds, upstream's pruned shape, inner values = max of children) byte-identical
    to the oracle's real-tree serialisation; read() restores every voxel's float32 value; round trips are identities."""
    rng = np.random.default_rng(61)
    t, r = octomap.OcTree(0.1), oo.OcTree(0.1)
    assert t.write() == r.write_ot_bytes()                                   # empty tree
    origin = np.array([0.1, 0.2, -0.1])
    for s in range(3):
        sc = _scan(rng, 5000, origin + 0.2 * s, far=15.0).astype(np.float32)
        t.insertPointCloud(sc, origin + 0.2 * s, maxrange=10.0)
        r.insertPointCloud_f32(sc, origin + 0.2 * s, 10.0)
    g = (np.arange(16) + 0.5) * 0.1 + 20.0                                    # a solid block: prunes up to depth 12
    blk = np.stack(np.meshgrid(g, g, g, indexing="ij"), axis=-1).reshape(-1, 3)
    for _ in range(5):                                                        # saturate: equal values -> collapsible
        t.updateNodes(blk, True)
        r.updateNodes(blk, True)
    pts = rng.normal(scale=1.0, size=(3000, 3))
    t.updateNodes(pts, False)
    r.updateNodes(pts, False)
    ot = t.write()
    assert ot == r.write_ot_bytes()
    assert t.size() == r.size() == (len(ot.split(b"data\n", 1)[1]) // 5)
    p = tmp_path / "a.ot"
    assert t.write(bytes(str(p), "utf-8")) and p.read_bytes() == ot
    u = octomap.OcTree(0.3)
    assert u.read(bytes(str(p), "utf-8")) and u.getResolution() == 0.1
    k0, v0 = t.voxels()
    k1, v1 = u.voxels()
    assert np.array_equal(k0, k1) and np.array_equal(v0.view(np.uint32), v1.view(np.uint32))
    assert u.write() == ot and u.writeBinary() == t.writeBinary()
    w = octomap.OcTree(0.1)
    assert w.read(ot) and w.write() == ot
    # .bt and .ot are not interchangeable, and truncated payloads are rejected
    with pytest.raises(Exception):
        octomap.OcTree(0.1).read(t.writeBinary())
    with pytest.raises(Exception):
        octomap.OcTree(0.1).readBinary(ot)
    with pytest.raises(Exception):
        octomap.OcTree(0.1).read(ot[:len(ot) - 7])
    assert not octomap.OcTree(0.1).read(bytes(str(tmp_path / "missing.ot"), "utf-8"))


def test_sequence_to_octree_modes(octomap, r3d):
    """mapping.sequence_to_octree: frames -> fused K1 -> one insertPointCloud per frame, with and without dropping the
    invalid (Z = 0) pixels, against the oracle fed with the oracle's own points."""
    mapping = importlib.import_module("3d_reconstruction_system_b200.mapping")
    rng = np.random.default_rng(71)
    n, H, W = 3, 24, 40
    depth = rng.integers(0, 1500, size=(n, H, W)).astype(np.uint16)
    depth[rng.random(size=depth.shape) < 0.3] = 0
    q = np.stack([po.synth_pose(k, n)[0] for k in range(n)])
    t = np.stack([po.synth_pose(k, n)[1] for k in range(n)])
    intr = (60.0, 60.0, 20.0, 12.0)
    for drop in (False, True):
        tree = mapping.sequence_to_octree(depth, q, t, intr, resolution=0.1, maxrange=6.0, depth_scale=1 / 256.0, drop_invalid=drop)
        ref = oo.OcTree(0.1)
        for k in range(n):
            rinv = po.quat_to_rinv_fixed(q[k])
            w = po.depth_to_world(depth[k], intr, rinv, t[k], po.MODE_DEPTH, 1 / 256.0)[1].astype(np.float32)
            if drop:
                w = w[po.valid_mask(depth[k], 0, 1 / 256.0).ravel()]
            ref.insertPointCloud_f32(w, po.camera_centre(rinv, t[k]), 6.0)
        assert_same_tree(tree, ref)


def test_pipelined_batch_record_overflow_path(octomap, r3d):
    """Scans whose deltas hold far more than the 65 536 records the pipeline's buffers start with: the emit kernel raises
    the device-side abort flag, the queued next scan skips, the host grows the buffers, lists again and re-queues.
    The result must still equal the scan-by-scan loop (whose parity with the oracle the other tests establish; the
    oracle itself needs minutes for rays this long)."""
    ctx = r3d.default_context(0)
    rng = np.random.default_rng(83)
    S, N = 3, 8000
    origins = np.array([[0.5 * s, 0.0, 0.0] for s in range(S)])
    scans = []
    for s in range(S):
        d = rng.normal(size=(N, 3))
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        scans.append((origins[s] + d * rng.uniform(60.0, 100.0, size=(N, 1))).astype(np.float32))
    scans = np.stack(scans)
    dev = ctx.to_device(scans.reshape(-1, 3))
    a, b = octomap.OcTree(0.1), octomap.OcTree(0.1)
    b.insertPointClouds(dev, origins, maxrange=90.0)              # fresh tree: buffers at their initial size
    assert b.lastScanStats()["records"] > 65536
    for s in range(S):
        a.insertPointCloud(scans[s], origins[s], maxrange=90.0)
    ka, va = a.voxels()
    kb, vb = b.voxels()
    assert np.array_equal(ka, kb) and np.array_equal(va.view(np.uint32), vb.view(np.uint32))
    assert a.writeBinary() == b.writeBinary()


def test_pool_growth_in_place_and_malloc_fallback_agree(octomap, r3d, tmp_path):
    """The brick pool grows in place (virtual memory management: chunks mapped behind a reserved address range) while scans are
    in flight; R3D_POOL_MALLOC=1 keeps the allocate-copy-free pool.  Both must give the oracle's tree, through several growth
    steps of the pool and re-hashes of the table inside one pipelined call."""
    import hashlib
    import os
    import subprocess
    import sys
    rng = np.random.default_rng(97)
    S, N = 12, 6000
    origins = np.array([[3.0 * s, 0.5 * s, 0.0] for s in range(S)])          # the sensor moves: new bricks with every scan
    scans = np.stack([_scan(rng, N, origins[s], far=25.0) for s in range(S)]).astype(np.float32)
    ctx = r3d.default_context(0)
    dev = ctx.to_device(scans.reshape(-1, 3))
    t, ref = octomap.OcTree(0.05), oo.OcTree(0.05)
    t.insertPointClouds(dev, origins, maxrange=20.0)
    for s in range(S):
        ref.insertPointCloud_f32(scans[s], origins[s], 20.0)
    g = t.growthStats()
    assert g["pool_regrowths"] >= 2 and g["table_regrowths"] >= 2, g        # the tree started at its minimum size
    assert_same_tree(t, ref)
    want = hashlib.sha256(t.writeBinary()).hexdigest()
    np.save(tmp_path / "scans.npy", scans)
    np.save(tmp_path / "origins.npy", origins)
    code = ("import importlib,sys,hashlib,numpy as np\n"
            "sys.path.insert(0, %r)\n"
            "r3d = importlib.import_module('3d_reconstruction_system_b200'); om = importlib.import_module('3d_reconstruction_system_b200.octomap')\n"
            "ctx = r3d.default_context(0); sc = np.load(%r); org = np.load(%r)\n"
            "t = om.OcTree(0.05); t.insertPointClouds(ctx.to_device(sc.reshape(-1, 3)), org, maxrange=20.0)\n"
            "print(hashlib.sha256(t.writeBinary()).hexdigest(), t.growthStats()['pool_regrowths'])\n") % (
                os.path.dirname(os.path.dirname(os.path.abspath(__file__))), str(tmp_path / "scans.npy"), str(tmp_path / "origins.npy"))
    p = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, R3D_POOL_MALLOC="1"), capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout + p.stderr
    got, grows = p.stdout.split()[-2:]
    assert got == want and int(grows) >= 2
