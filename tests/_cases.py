"""Shared adversarial inputs for the text-formatting tests."""
import numpy as np


def fixed4_cases(seed=11, n=20000):
    """Doubles that stress "%.4f": ties on the 5th decimal (k/32, k/64 ...), values around powers of ten, denormals, huge
    magnitudes (bignum path), signed zeros, values that round to +-0.0000, inf and nan."""
    rng = np.random.default_rng(seed)
    v = [rng.normal(scale=s, size=n) for s in (1e-5, 1e-3, 1.0, 50.0, 3000.0, 1e7, 1e12)]
    v.append(rng.integers(-640000, 640000, size=n) / 32.0)              # exact ties: x.xxxx5 with finite binary expansion
    v.append(rng.integers(-10**6, 10**6, size=n) / 64.0 + rng.integers(0, 2, size=n) * 2.0 ** -20)
    v.append(10.0 ** rng.integers(-8, 18, size=n) * rng.choice([1.0, 0.99999999999, 1.00000000001, 9.99995, 0.99995], size=n))
    v.append(np.ldexp(rng.uniform(0.5, 1.0, size=n), rng.integers(-1074, 1024, size=n)) * rng.choice([-1.0, 1.0], size=n))
    v.append(np.array([0.0, -0.0, 5e-5, -5e-5, 4.9999999e-5, 0.00005000000000000001, 0.5, 1e-4, 9999.99995, 9.2e14, 9.3e14, 2.0 ** 63, 1.7976931348623157e308,
                       5e-324, -5e-324, np.inf, -np.inf, np.nan, 0.03125, 0.09375, 1234.56785, -0.00004, 99999.99995, 0.99995]))
    return np.concatenate(v)


def repr_cases(seed=12, n=200000):
    """Doubles that stress the shortest-repr conversion: uniformly random bit patterns (every exponent), decimal literals
    with few digits (trailing-zero paths), integers up to 2^63, powers of two and ten and their neighbours, subnormals,
    the repr() layout thresholds 1e-4 / 1e16, and the back-projection's own value range."""
    rng = np.random.default_rng(seed)
    bits = rng.integers(0, 1 << 63, size=n, dtype=np.uint64) | (rng.integers(0, 2, size=n, dtype=np.uint64) << np.uint64(63))
    v = [bits.view(np.float64)]
    v.append(np.round(rng.normal(scale=100.0, size=n), rng.integers(0, 8)))
    v.append(rng.integers(-10**6, 10**6, size=n) / 10.0 ** rng.integers(0, 12, size=n))
    v.append(rng.integers(0, 1 << 62, size=n, dtype=np.int64).astype(np.float64))
    v.append(np.ldexp(1.0, rng.integers(-1074, 1024, size=n // 10)))
    p10 = 10.0 ** rng.integers(-320, 309, size=n // 10)
    v.append(np.concatenate([p10, np.nextafter(p10, np.inf), np.nextafter(p10, -np.inf)]))
    v.append(np.ldexp(rng.integers(1, 1 << 52, size=n // 10).astype(np.float64), -1074))
    v.append((rng.integers(0, 1242, size=n) - 607.1928) / 718.856 * rng.integers(0, 65536, size=n))
    v.append(np.array([0.0, -0.0, 1e-4, 9.999999999999999e-05, 1e-5, 1e16, 9999999999999998.0, 1.2345678901234567e16, 1e22, 1e23, 5e-324, 2.2250738585072014e-308,
                       1.7976931348623157e308, 0.1, 0.2, 0.3, 1 / 3, 2 / 3, 123456.789, 1e15, 1e17, 4.35, 0.5, 1.0, -1.5, 9007199254740993.0, np.inf, -np.inf, np.nan, 2.0 ** 63]))
    out = np.concatenate(v)
    return out[np.isfinite(out) | (np.arange(out.size) >= out.size - 30)]


def load_c1(golden_dir):
    """BASELINE config 1 fixture (oracle/gen_golden.py section 5): (npz, meta, the 16-bit frame regenerated and checked)."""
    import hashlib
    import json
    import os
    from oracle import points_oracle as po
    g = np.load(os.path.join(golden_dir, "ref_c1_kitti.npz"))
    with open(os.path.join(golden_dir, "ref_meta.json")) as f:
        meta = json.load(f)["c1_kitti"]
    d16 = po.synth_depth_u16(meta["W"], meta["H"], po.KITTI_INTRINSICS, meta["seed"], "street")
    assert hashlib.sha256(d16.tobytes()).hexdigest() == meta["depth16_sha256"]          # the generator still makes the same frame
    assert np.array_equal((d16 >> 8).astype(np.uint8), g["depth8"])                      # what IMREAD_GRAYSCALE hands the reference
    return g, meta, d16
