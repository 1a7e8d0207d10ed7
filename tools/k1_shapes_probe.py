"""Ad-hoc probe: the bulk K1 kernels on narrow / short images (many row and frame wraps per 1024-pixel group)."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import points_oracle as po
r3d = importlib.import_module("3d_reconstruction_system_b200")
ctx = r3d.Context(0)
rng = np.random.default_rng(5)
bad = 0
for shape in [(6, 8, 256), (5, 9, 257), (3, 40, 300), (3, 12, 1000), (2, 8, 4097), (40, 8, 258)]:
    n, H, W = shape
    for dtype, scale in ((np.uint16, 1 / 256.0), (np.float32, 1.0)):
        d = (rng.integers(0, 60000, size=shape) * (rng.random(shape) > 0.2)).astype(dtype)
        q = rng.normal(size=(n, 4)); t = rng.normal(size=(n, 3)) * 10
        rt = ctx.pose_to_rt(q, t)
        ref = np.concatenate([po.depth_to_world(d[k], po.REF_INTRINSICS, rt[k, :9].reshape(3, 3), rt[k, 9:], 0, scale)[1] for k in range(n)])
        mask = np.concatenate([po.valid_mask(d[k], 0, scale).ravel() for k in range(n)])
        for od in (np.float32, np.float64):
            got, _ = ctx.backproject(d, po.REF_INTRINSICS, rt=rt, depth_scale=scale, out_dtype=od)
            ok1 = np.array_equal(got, ref.astype(od))
            gc, cnt = ctx.backproject(d, po.REF_INTRINSICS, rt=rt, depth_scale=scale, compact=True, out_dtype=od)
            ok2 = np.array_equal(gc, ref[mask].astype(od)) and cnt.tolist() == [int(po.valid_mask(d[k], 0, scale).sum()) for k in range(n)]
            if not (ok1 and ok2):
                bad += 1
                print("MISMATCH", shape, dtype.__name__, od.__name__, ok1, ok2)
print("shapes probe:", "all bit-identical" if bad == 0 else "%d mismatches" % bad)
