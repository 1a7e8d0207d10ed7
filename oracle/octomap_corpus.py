"""Deterministic replay corpus for the occupancy half of the path (a9-a13) -- TEST INFRASTRUCTURE ONLY.

One list of cases, two consumers:
  * oracle/pin_against_octomap.py replays it against the REAL `octomap` extension (the un-vendored dependency of
    octomap/txt_transfer_octomap.py:2,25,33-36 and octomap/ply_transfer_octomap.py:2,33,45-48) wherever that module can
    be imported, and writes the answers to tests/golden/octomap_pin.json;
  * tests/test_oracle_octomap.py replays it against oracle/octomap_oracle.c and compares with that file when it exists.

A case is a dict: name, res, ops (a list of operations on ONE tree), probes (coordinates whose log-odds are read at the
end).  Operations use only calls both upstream bindings offer with the same meaning:
  ("update", points (n,3) float64, True|False|float)        -> tree.updateNode(p, v) per point, in order
  ("insert", points (n,3) float64, origin (3,), maxrange, discretize) -> tree.insertPointCloud(...)
What is recorded per case: size(), sha256 + length of writeBinary() (called last: it mutates the tree), and the float32
bit pattern of getLogOdds() at every probe (None when the voxel is unknown), read BEFORE writeBinary.
Everything is generated from fixed seeds; nothing here reads files.
"""
import numpy as np

from . import points_oracle as po

SEED = 20261018


def _cube(n, step, base):
    g = np.stack(np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij"), -1).reshape(-1, 3)
    return g * step + step / 2 + np.asarray(base, dtype=np.float64)


def kitti_scan(k, n_total=4500, stride=1):
    """Scan k of the synthetic C2/C3 sequence as float32 world points + sensor origin (what K1 hands to insertPointCloud)."""
    W, H = 1242, 375
    d16 = po.synth_depth_u16(W, H, po.KITTI_INTRINSICS, SEED + 2 + (k % 32), "street")
    q, t = po.synth_pose(k, n_total)
    rinv = po.quat_to_rinv_fixed(q)
    world = po.depth_to_world(d16, po.KITTI_INTRINSICS, rinv, t, po.MODE_DEPTH, 1.0 / 256.0)[1].astype(np.float32)
    return world[::stride].astype(np.float64), po.camera_centre(rinv, t).astype(np.float32).astype(np.float64)


def cases(heavy=True):
    rng = np.random.default_rng(SEED)
    out = []

    def add(name, res, ops, probes=()):
        out.append({"name": name, "res": res, "ops": ops, "probes": np.asarray(probes, dtype=np.float64).reshape(-1, 3)})

    # ---- coordToKey boundaries (a10): in/out of range, negative zero, float32 rounding of 0.3 / 0.7, key-space edges
    edge = np.array([[0, 0, 0], [-1e-9, 0.05, 0.1], [0.3, 0.7, -0.3], [3276.75, 0, 0], [3276.8, 0, 0], [-3276.8, 0, 0], [-3276.75, 0, 0],
                     [0.1, 0.2, 0.30000001], [1e30, 0, 0], [-0.0, -0.0, -0.0], [3276.79, -3276.79, 3276.79]], dtype=np.float64)
    grid = np.arange(-40, 40)[:, None] * 0.1 + np.array([0.0, 1e-7, -1e-7])[None, :]
    edge = np.concatenate([edge, np.stack([grid.ravel(), grid.ravel()[::-1], grid.ravel()], axis=1)])
    add("key_boundaries_0.1", 0.1, [("update", edge, True)], edge[:40])
    add("key_boundaries_0.05", 0.05, [("update", np.concatenate([edge / 2, [[1638.39, 0, 0], [1638.41, 0, 0]]]), True)], edge[:40] / 2)

    # ---- clamped log-odds ladder, early abort, float-typed update (a10)
    p = np.array([[1.0, 2.0, 3.0]])
    ops = [("update", p, True)] * 7 + [("update", p, False)] * 20 + [("update", p, 1.0), ("update", p, -0.25), ("update", p, True)]
    add("ladder", 0.1, ops, p)
    for k in range(1, 7):
        add("ladder_hits_%d" % k, 0.1, [("update", p, True)] * k, p)
    add("ladder_misses_6", 0.1, [("update", p, False)] * 6, p)

    # ---- .bt structure (a13): single voxels, siblings pruned at update time, mixed children, prune()'s early-break quirk
    add("bt_single_occ", 0.1, [("update", np.array([[0.05, 0.05, 0.05]]), True)])
    add("bt_single_free_0.05", 0.05, [("update", np.array([[-0.01, -0.01, -0.01]]), False)])
    add("bt_eight_siblings", 0.1, [("update", _cube(2, 0.1, [0, 0, 0]), True)])
    add("bt_mixed_children", 0.1, [("update", np.array([[0.05, 0.05, 0.05], [0.05, 0.15, 0.15]]), True), ("update", np.array([[0.15, 0.05, 0.05]]), False)])
    c4 = _cube(4, 0.1, [0, 0, 0])
    twice = c4[~((c4[:, 0] < 0.2) & (c4[:, 1] < 0.2) & (c4[:, 2] < 0.2))]
    add("bt_prune_quirk_a", 0.1, [("update", c4, True), ("update", twice, True)])
    add("bt_prune_quirk_b", 0.1, [("update", c4, True), ("update", c4[(np.round(c4[:, 0] * 10 - 0.5).astype(int) % 2) == 0], True)])
    add("bt_saturated_block_16", 0.1, [("update", _cube(16, 0.1, [10, -3, 2]), True)] * 6, _cube(2, 0.8, [10, -3, 2]))
    add("bt_res_text_0.25", 0.25, [("update", np.array([[0.3, -0.3, 7.0]]), True)])
    add("bt_res_text_0.033", 0.033, [("update", np.array([[0.3, -0.3, 7.0]]), True)])

    # ---- computeRayKeys edge cases (a11), one ray per insertPointCloud so that the free set IS the ray
    o = np.array([0.05, 0.05, 0.05])
    rays = {"axis_x": [1.05, 0.05, 0.05], "axis_-y": [0.05, -0.95, 0.05], "axis_z": [0.05, 0.05, 2.05], "same_voxel": [0.06, 0.06, 0.06],
            "diag_xy_tie": [1.05, 1.05, 0.05], "diag_xyz_tie": [1.05, 1.05, 1.05], "diag_-x-y-z_tie": [-0.95, -0.95, -0.95],
            "diag_xz_tie": [2.05, 0.05, 2.05], "diag_yz_tie": [0.05, -1.95, 1.05], "near_tie": [1.05, 1.0500001, 0.05],
            "shallow": [5.05, 0.15, 0.05], "steep": [0.15, 0.25, 5.05], "long": [60.0, -35.0, 12.5], "out_of_bounds_end": [4000.0, 0.0, 0.0]}
    for name, e in rays.items():
        add("ray_" + name, 0.1, [("insert", np.array([e]), o, -1.0, False)], [e, o])
    add("ray_origin_on_border", 0.1, [("insert", np.array([[1.0, 1.0, 1.0]]), np.array([0.0, 0.0, 0.0]), -1.0, False)])
    add("ray_origin_negative_border", 0.1, [("insert", np.array([[-1.0, 2.0, -3.0]]), np.array([-0.1, 0.2, -0.3]), -1.0, False)])
    add("ray_origin_out_of_bounds", 0.1, [("insert", np.array([[1.0, 1.0, 1.0]]), np.array([5000.0, 0.0, 0.0]), -1.0, False)])
    add("ray_maxrange_truncates", 0.1, [("insert", np.array([[1.05, 0.05, 0.05], [2.05, 0.05, 0.05], [100.0, 0.05, 0.05]]), o, 10.0, False)],
        [[1.05, 0.05, 0.05], [0.25, 0.05, 0.05], [10.05, 0.05, 0.05], [10.15, 0.05, 0.05]])
    add("ray_maxrange_exact", 0.1, [("insert", np.array([[10.05, 0.05, 0.05], [0.05, 10.05, 0.05]]), o, 10.0, False)])
    add("ray_0.05_diag", 0.05, [("insert", np.array([[1.025, 1.025, -1.025]]), np.array([0.025, 0.025, -0.025]), -1.0, False)])

    # ---- computeUpdate set semantics: occupied wins over free inside one scan, every key once per scan, scans accumulate
    fan = np.array([[3.0 * np.cos(a), 3.0 * np.sin(a), 0.3 * np.sin(5 * a)] for a in np.linspace(0, 2 * np.pi, 400, endpoint=False)])
    add("scan_fan_occupied_wins", 0.1, [("insert", np.concatenate([fan, fan * 0.5]), np.zeros(3), -1.0, False)], fan[::40] * 0.5)
    add("scan_fan_three_times", 0.1, [("insert", fan, np.zeros(3), -1.0, False)] * 3, fan[::40])
    add("scan_fan_shifted_origins", 0.1, [("insert", fan + np.array([0.03 * i, 0, 0]), np.array([0.03 * i, 0.0, 0.0]), 2.5, False) for i in range(8)], fan[::40])

    # ---- random clouds: both resolutions, bounded / unbounded range, discretize, far from the origin
    for i, (res, spread, mr, disc, centre) in enumerate([(0.1, 4.0, -1.0, False, (0, 0, 0)), (0.1, 15.0, 10.0, False, (0, 0, 0)), (0.05, 3.0, -1.0, False, (0, 0, 0)),
                                                          (0.05, 6.0, 4.0, False, (100.3, -50.7, 3.1)), (0.1, 6.0, -1.0, True, (0, 0, 0)),
                                                          (0.1, 8.0, 6.0, True, (-1200.5, 7.25, 2900.0)), (0.1, 30.0, 80.0, False, (1500.25, -3.5, 1720.125))]):
        c = np.asarray(centre, dtype=np.float64)
        pts = rng.normal(scale=spread, size=(3000, 3)) + c
        add("random_cloud_%d" % i, res, [("insert", pts, c + np.array([0.01, -0.02, 0.03]), mr, disc)], pts[:16])
    two = rng.normal(scale=5.0, size=(2, 2000, 3))
    add("random_two_scans_then_updates", 0.1, [("insert", two[0], np.zeros(3), 12.0, False), ("insert", two[1], np.array([1.0, 0.5, -0.2]), 12.0, False),
                                               ("update", two[0][:500], True), ("update", two[1][:500], False)], two[0][:16])

    if heavy:
        # ---- the benchmark's own scans (config 3): KITTI-shape street frames at 0.1 m / 80 m; Z = 0 sky pixels are rays to the origin
        s0 = kitti_scan(2250)
        add("kitti_scan_2250", 0.1, [("insert", s0[0], s0[1], 80.0, False)], s0[0][::46575])
        add("kitti_scans_2250_2252", 0.1, [("insert",) + kitti_scan(2250 + i) + (80.0, False) for i in range(3)], s0[0][::46575])
        s1 = kitti_scan(2250, stride=7)
        add("kitti_scan_2250_0.05_stride7", 0.05, [("insert", s1[0], s1[1], 80.0, False)], s1[0][::6654])
        add("kitti_scan_first_frame", 0.1, [("insert",) + kitti_scan(0, stride=3) + (80.0, False)])
        add("kitti_scan_updatenode_mode", 0.1, [("update", kitti_scan(2250, stride=5)[0], True)])
    return out
