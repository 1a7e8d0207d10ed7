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
