"""The kernels' scalar arithmetic (csrc/r3d_math.cuh) is host/device code; here it is compiled for the host and
checked against the oracle, so formula errors surface without a GPU.  (The GPU parity tests check the kernels.)"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import octomap_oracle as oo
from oracle import points_oracle as po
from _cases import fixed4_cases, repr_cases

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def hm():
    out = os.path.join(HERE, "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libhostmath.so")
    src = os.path.join(HERE, "hostmath", "hostmath.cpp")
    hdr = os.path.join(HERE, "..", "3d_reconstruction_system_b200", "csrc", "r3d_math.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-x", "c++", src, "-o", so])
    L = C.CDLL(so)
    L.hm_ray_keys.restype = C.c_long
    L.hm_ray_keys.argtypes = [C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long]
    L.hm_ray_keys_predicated.restype = C.c_long
    L.hm_ray_keys_predicated.argtypes = [C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long]
    L.hm_ray_keys_pending.restype = C.c_long
    L.hm_ray_keys_pending.argtypes = [C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long]
    L.hm_coord_to_key.argtypes = [C.c_double, C.c_float, C.c_float, C.c_float, C.c_void_p]
    L.hm_scan_point_end.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
    L.hm_pose_apply.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.hm_pixel_coeff.restype = C.c_double
    L.hm_pixel_coeff.argtypes = [C.c_int, C.c_double, C.c_double]
    L.hm_clamped_add.restype = C.c_float
    L.hm_clamped_add.argtypes = [C.c_float] * 4
    L.hm_brick_key.restype = C.c_uint64
    L.hm_brick_key.argtypes = [C.c_uint32] * 3
    L.hm_brick_morton.restype = C.c_uint64
    L.hm_brick_morton.argtypes = [C.c_uint64]
    L.hm_brick_voxel_index.restype = C.c_uint32
    L.hm_brick_voxel_index.argtypes = [C.c_uint32] * 3
    L.hm_brick_voxel_coords.argtypes = [C.c_uint32, C.c_void_p]
    L.hm_txt_rows.restype = C.c_long
    L.hm_txt_rows.argtypes = [C.c_void_p, C.c_long, C.c_int, C.c_void_p]
    L.hm_fixed4_units.restype = C.c_long
    L.hm_fixed4_units.argtypes = [C.c_void_p, C.c_long, C.c_void_p]
    L.hm_ply_rows_fast.restype = C.c_long
    L.hm_ply_rows_fast.argtypes = [C.c_void_p, C.c_long, C.c_void_p, C.c_void_p]
    L.hm_ply_rows.restype = C.c_long
    L.hm_ply_rows.argtypes = [C.c_void_p, C.c_long, C.c_void_p, C.c_void_p]
    return L


def test_fixed4_rows_match_python_percent_format(hm):
    x = fixed4_cases()
    x = x[: (x.size // 3) * 3].reshape(-1, 3).copy()
    out = C.create_string_buffer(x.shape[0] * 3 * 330 + 64)
    n = hm.hm_ply_rows(x.ctypes.data, x.shape[0], None, out)
    want = "".join("%.4f %.4f %.4f \n" % (a, b, c) for a, b, c in x.tolist())
    assert out.raw[:n].decode("ascii") == want
    rgb = np.random.default_rng(5).integers(0, 256, size=(x.shape[0], 3), dtype=np.uint8)
    rgb[:3] = [[0, 9, 10], [99, 100, 255], [1, 2, 3]]
    n = hm.hm_ply_rows(x.ctypes.data, x.shape[0], rgb.ctypes.data, out)
    want = "".join("%.4f %.4f %.4f %d %d %d 0\n" % (a, b, c, r, g, bb) for (a, b, c), (r, g, bb) in zip(x.tolist(), rgb.tolist()))
    assert out.raw[:n].decode("ascii") == want


def test_fixed4_units_by_fma_equal_the_integer_rule(hm):
    """K6's fast path (one fma, r3d_math.cuh::fixed4_units_fma) against the exact integer rounding it replaces: the
    adversarial set, every tie k/32 and k/64 in range, random doubles of every exponent below the limit, the limit's
    neighbours."""
    rng = np.random.default_rng(77)
    ties = np.arange(-429496 * 32, 429496 * 32, 7, dtype=np.float64) / 32.0
    near = ties[::5] + np.ldexp(rng.choice([-1.0, 1.0], size=ties[::5].size), rng.integers(-60, -30, size=ties[::5].size))
    spread = np.ldexp(rng.uniform(0.5, 1.0, size=400000), rng.integers(-1074, 19, size=400000)) * rng.choice([-1.0, 1.0], size=400000)
    edge = np.array([429495.99994999997, 429495.99995, 429495.9999, np.nextafter(429496.0, 0.0), 429496.0, -429495.99995, 0.0, -0.0, 5e-5, 4.9999999999999996e-5,
                     5.000000000000001e-5, 1.5e-4, 2.5e-4, 0.00015, 0.00025, 0.00035])
    v = np.concatenate([fixed4_cases(), ties, near, spread, edge])
    bad = C.c_long(0)
    n = hm.hm_fixed4_units(v.ctypes.data, v.size, C.byref(bad))
    assert n > 1_000_000 and bad.value == 0


def test_fixed4_fast_path_rows_match_python(hm):
    """The row writer of K6 (fast4_measure / fast4_write: 32-bit digits, every length from 1 to 6 digits before the point,
    signs, values that round up into the next length) on the host, against Python's "%.4f"."""
    rng = np.random.default_rng(78)
    mags = np.concatenate([10.0 ** np.arange(-6, 6), 10.0 ** np.arange(0, 6) - 5e-5, 10.0 ** np.arange(0, 6) - 5.1e-5, [429495.9999, 99999.99995, 9.99995]])
    edge = np.concatenate([mags, -mags, [0.0, -0.0, -4e-5, 4e-5]])
    v = np.concatenate([edge, rng.uniform(-429000, 429000, 300000), rng.normal(scale=30.0, size=300000), rng.normal(scale=0.01, size=100000),
                        rng.integers(-10**7, 10**7, size=100000) / 32.0, fixed4_cases()])
    v = v[: (v.size // 3) * 3].reshape(-1, 3).copy()
    out = C.create_string_buffer(v.shape[0] * 3 * 330 + 64)
    fast = C.c_long(0)
    n = hm.hm_ply_rows_fast(v.ctypes.data, v.shape[0], out, C.byref(fast))
    want = "".join("%.4f %.4f %.4f \n" % (a, b, c) for a, b, c in v.tolist())
    assert out.raw[:n].decode("ascii") == want
    assert fast.value > 200000


def ray(hm, res, o, e):
    o32, e32 = np.asarray(o, np.float32), np.asarray(e, np.float32)
    buf = np.zeros((300000, 3), np.uint16)
    n = hm.hm_ray_keys(res, o32.ctypes.data, e32.ctypes.data, buf.ctypes.data, buf.shape[0])
    # the kernel's branch-free form of the same walk must list the same keys
    buf2 = np.zeros((300000, 3), np.uint16)
    n2 = hm.hm_ray_keys_predicated(res, o32.ctypes.data, e32.ctypes.data, buf2.ctypes.data, buf2.shape[0])
    assert n2 == n and (n < 0 or np.array_equal(buf[:n], buf2[:n]))
    n3 = hm.hm_ray_keys_pending(res, o32.ctypes.data, e32.ctypes.data, buf2.ctypes.data, buf2.shape[0])
    assert n3 == n and (n < 0 or np.array_equal(buf[:n], buf2[:n]))
    return None if n < 0 else buf[:n].copy()


@pytest.mark.parametrize("res", [0.1, 0.05, 0.25])
def test_dda_matches_oracle(hm, res):
    rng = np.random.default_rng(3)
    t = oo.OcTree(res)
    cases = []
    for _ in range(1500):
        o = rng.uniform(-20, 20, 3)
        cases.append((o, o + rng.normal(size=3) * rng.uniform(0, 60)))
    o = np.array([0.05, 0.05, 0.05])
    cases += [(o, o), (o, o + [3, 0, 0]), (o, o + [0, -3, 0]), (o, o + [0, 0, 3]), (o, o + [1, 1, 1]), (o, o + [1e-3, 0, 0]),
              (o, [5000, 0, 0]), ([-5000, 0, 0], o), (o, o + [res, res, 0]), (np.zeros(3), [res * 10, res * 10, res * 10]),
              (np.zeros(3), [-res * 7, res * 7, 0]), ([1600.0, -1600.0, 3.0], [1630.0, -1570.0, 9.0])]
    for o, e in cases:
        got = ray(hm, res, o, e)
        want = t.computeRayKeys(np.float32(o), np.float32(e), 300000)
        if want is None:
            assert got is None
        else:
            assert np.array_equal(got, want), (o, e)


def test_keys_and_maxrange_end(hm):
    rng = np.random.default_rng(4)
    t = oo.OcTree(0.1)
    pts = np.concatenate([rng.uniform(-3300, 3300, size=(5000, 3)), (np.arange(-50, 50)[:, None] * 0.1 + np.zeros((1, 3)))])
    k = (C.c_uint16 * 3)()
    for p in pts.astype(np.float32):
        ok = hm.hm_coord_to_key(10.0, float(p[0]), float(p[1]), float(p[2]), k)
        want = t.coordToKey(p.astype(np.float64))
        assert bool(ok) == (want is not None)
        if want:
            assert (k[0], k[1], k[2]) == want
    # truncated ray end == the oracle's (checked through the free-key set of a single beyond-range point)
    o = np.array([0.3, -0.2, 0.1], np.float32)
    for _ in range(300):
        p = (o + rng.normal(size=3) * 40).astype(np.float32)
        e = np.zeros(3, np.float32)
        inr = hm.hm_scan_point_end(o.ctypes.data, p.ctypes.data, 15.0, e.ctypes.data)
        fr, oc = t.computeUpdate(p[None, :], o, 15.0)
        assert bool(inr) == (len(oc) == 1)
        got = ray(hm, 0.1, o, e)
        keys = oo.pack_keys(got) if got is not None and len(got) else np.zeros(0, np.uint64)
        if inr and len(oc):
            keys = keys[keys != oc[0]]
        assert np.array_equal(np.sort(np.unique(keys)), fr)


def test_pose_tables_and_clamp(hm, golden_dir):
    z = np.load(os.path.join(golden_dir, "ref_c2w_small.npz"))
    for i in (0, 5, 319, 320, 321, 1241):
        assert hm.hm_pixel_coeff(i, 320.0, 600.391) == (i - 320) / 600.391
        assert hm.hm_pixel_coeff(i, 607.1928, 718.856) == (i - 607.1928) / 718.856
    k = 2
    cam, world = po.depth_to_world(z["depths"][k], po.REF_INTRINSICS, z["rinv"][k], z["trans"][k])
    rt = np.concatenate([z["rinv"][k].reshape(9), z["trans"][k]])
    w = np.zeros(3)
    for i in range(cam.shape[0]):
        p = np.ascontiguousarray(cam[i])
        hm.hm_pose_apply(rt.ctypes.data, p.ctypes.data, w.ctypes.data)
        assert np.array_equal(w, world[i])
    v = np.float32(0)
    for _ in range(8):
        v = np.float32(hm.hm_clamped_add(v, np.float32(0.84729785), np.float32(-2.0000279), np.float32(3.5110307)))
    assert v == np.float32(3.5110307)


def test_brick_indexing_roundtrip(hm):
    seen = set()
    xyz = (C.c_uint32 * 3)()
    for x in range(8):
        for y in range(8):
            for z in range(8):
                i = hm.hm_brick_voxel_index(x + 8 * 5, y + 8 * 9, z)
                assert i == (x & 1) | ((y & 1) << 1) | ((z & 1) << 2) | ((x >> 1 & 1) << 3) | ((y >> 1 & 1) << 4) | ((z >> 1 & 1) << 5) | \
                    ((x >> 2) << 6) | ((y >> 2) << 7) | ((z >> 2) << 8)
                hm.hm_brick_voxel_coords(i, xyz)
                assert (xyz[0], xyz[1], xyz[2]) == (x, y, z)
                seen.add(i)
    assert len(seen) == 512
    bk = hm.hm_brick_key(65535, 8, 32768)
    assert bk == (8191 | (1 << 13) | (4096 << 26))
    m = hm.hm_brick_morton(bk)
    want = 0
    for b in range(13):
        want |= ((8191 >> b) & 1) << (3 * b) | ((1 >> b) & 1) << (3 * b + 1) | ((4096 >> b) & 1) << (3 * b + 2)
    assert m == want


def test_shortest_repr_rows_match_python_repr(hm):
    x = repr_cases()
    x = x[: (x.size // 3) * 3].reshape(-1, 3).copy()
    out = C.create_string_buffer(x.shape[0] * 76 + 64)
    n = hm.hm_txt_rows(x.ctypes.data, x.shape[0], 0, out)
    want = "".join("%r,%r,%r\n" % (a, b, c) for a, b, c in x.tolist())
    got = out.raw[:n].decode("ascii")
    if got != want:
        g, w = got.split("\n"), want.split("\n")
        bad = [(a, b) for a, b in zip(g, w) if a != b][:5]
        raise AssertionError("first mismatches (got, want): %r" % bad)
    # camera txt: Z is the raw integer sample
    y = x[:5000].copy()
    y[:, 2] = np.random.default_rng(1).integers(0, 65536, size=5000)
    n = hm.hm_txt_rows(y.ctypes.data, y.shape[0], 1, out)
    assert out.raw[:n].decode("ascii") == "".join("%r,%r,%d\n" % (a, b, int(c)) for a, b, c in y.tolist())


def test_axis_as_multiplier_update_is_exact():
    """The dense ray walker applies the chosen axis by multiplication (tMax + m * tDelta, m in {0.0, 1.0}; r3d_octree.cu,
    K3_VARIANT 1).  That equals "tMax + tDelta on the chosen axis, untouched elsewhere" bit for bit, including upstream's
    DBL_MAX sentinels for axes the ray does not move along."""
    import numpy as np
    rng = np.random.default_rng(9)
    tm = np.concatenate([rng.uniform(0, 1e3, 4096), [np.finfo(np.float64).max, 0.0, 5e-324, 1e-310]])
    td = np.concatenate([rng.uniform(1e-3, 1e3, 4096), [np.finfo(np.float64).max, np.finfo(np.float64).max, 5e-324, 0.1]])
    with np.errstate(over="ignore"):
        assert np.array_equal((tm + 1.0 * td).view(np.uint64), (tm + td).view(np.uint64))
        assert np.array_equal((tm + 0.0 * td).view(np.uint64), tm.view(np.uint64))
