"""CPU restatement (numpy float64) of the reference's depth -> camera -> world point path.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under 3d_reconstruction_system_b200/ or the
drop-in scripts may import it.

Pinned against the reference itself: oracle/gen_golden.py imports /root/reference/transfer/*.py
(with the three import shims SURVEY.md section 8c lists) in the authoring container and commits
inputs + outputs under tests/golden/; tests/test_oracle_points.py replays them here.

Each function cites the reference lines it follows (paths relative to /root/reference).
"""
import numpy as np

# Hard-coded intrinsics of the reference scripts (transfer/camera_to_world.py:68-71,
# transfer/pixel_to_camera.py:25-28).
REF_INTRINSICS = (600.391, 600.079, 320, 240)
# airsim/main.cpp:40-43
AIRSIM_INTRINSICS = (269.5, 269.5, 319.5, 239.5)
# KITTI odometry
KITTI_INTRINSICS = (718.856, 718.856, 607.1928, 185.2157)

MODE_DEPTH = 0
MODE_DISPARITY = 1


def raw_to_z(raw, mode=MODE_DEPTH, depth_scale=1.0, fB=0.0):
    """Depth decode to float64 Z.

    MODE_DEPTH: Z = raw * depth_scale.  depth_scale == 1.0 is the reference behaviour (Z is the raw
    pixel value, transfer/camera_to_world.py:75-78: `Z = Z`), the product by 1.0 being exact.
    MODE_DISPARITY (build-defined; not in the reference code):
    d = raw * depth_scale; Z = fB / d for d > 0 else 0 (invalid).
    """
    r = np.asarray(raw).astype(np.float64)
    d = r * np.float64(depth_scale)
    if mode == MODE_DEPTH:
        return d
    z = np.zeros_like(d)
    pos = d > 0
    z[pos] = np.float64(fB) / d[pos]
    return z


def backproject(z, intr=REF_INTRINSICS):
    """gentxtcord (transfer/camera_to_world.py:67-83, transfer/pixel_to_camera.py:24-44), vectorised:
    X = (i - cx)/fx*Z, Y = (j - cy)/fy*Z with i = column, j = row, evaluated ((i-cx)/fx)*Z.
    Returns X, Y, Z as (H, W) float64, row-major pixel order, every pixel emitted."""
    fx, fy, cx, cy = intr
    z = np.asarray(z, dtype=np.float64)
    H, W = z.shape
    a = (np.arange(W, dtype=np.float64) - np.float64(cx)) / np.float64(fx)
    b = (np.arange(H, dtype=np.float64) - np.float64(cy)) / np.float64(fy)
    X = a[None, :] * z
    Y = b[:, None] * z
    return X, Y, z


def scipy_transfer(quat):
    """transfer/camera_to_world.py:53-55 verbatim semantics: scipy scalar-last quaternion (normalised by
    scipy) -> rotation matrix -> general matrix inverse (np.matrix(...).I == np.linalg.inv)."""
    from scipy.spatial.transform import Rotation as R
    r = R.from_quat(np.asarray(quat, dtype=np.float64))
    return np.linalg.inv(np.asarray(r.as_matrix(), dtype=np.float64))


def quat_to_rinv_fixed(quat):
    """Library-free restatement of scipy_transfer with a fixed operation order (what the C-ABI's
    r3d_pose_to_rt computes): normalise, scipy's as_matrix formula, cofactor inverse.
    Agrees with scipy_transfer() to a few ulp(fp64); see tests/test_oracle_points.py."""
    q = np.asarray(quat, dtype=np.float64)
    n = np.sqrt(((q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]) + q[3] * q[3])
    if not n > 0:
        raise ValueError("Found zero norm quaternions in `quat`.")
    x, y, z, w = q[0] / n, q[1] / n, q[2] / n, q[3] / n
    x2, y2, z2, w2 = x * x, y * y, z * z, w * w
    xy, zw, xz, yw, yz, xw = x * y, z * w, x * z, y * w, y * z, x * w
    m = np.array([
        [((x2 - y2) - z2) + w2, 2.0 * (xy - zw), 2.0 * (xz + yw)],
        [2.0 * (xy + zw), ((-x2 + y2) - z2) + w2, 2.0 * (yz - xw)],
        [2.0 * (xz - yw), 2.0 * (yz + xw), ((-x2 - y2) + z2) + w2],
    ], dtype=np.float64)
    c00 = m[1, 1] * m[2, 2] - m[1, 2] * m[2, 1]
    c01 = m[1, 2] * m[2, 0] - m[1, 0] * m[2, 2]
    c02 = m[1, 0] * m[2, 1] - m[1, 1] * m[2, 0]
    det = (m[0, 0] * c00 + m[0, 1] * c01) + m[0, 2] * c02
    inv = np.array([
        [c00, m[0, 2] * m[2, 1] - m[0, 1] * m[2, 2], m[0, 1] * m[1, 2] - m[0, 2] * m[1, 1]],
        [c01, m[0, 0] * m[2, 2] - m[0, 2] * m[2, 0], m[0, 2] * m[1, 0] - m[0, 0] * m[1, 2]],
        [c02, m[0, 1] * m[2, 0] - m[0, 0] * m[2, 1], m[0, 0] * m[1, 1] - m[0, 1] * m[1, 0]],
    ], dtype=np.float64) / det
    return inv


def point_camera(P, r_inverse, t):
    """point_camera (transfer/camera_to_world.py:57-59): p_world = R^-1 . (p_cam - t), for P (..., 3).
    Fixed left-to-right evaluation, every product and sum rounded separately (numpy ufuncs never
    contract to FMA) -- the order the CUDA kernel mirrors with __dmul_rn/__dadd_rn."""
    P = np.asarray(P, dtype=np.float64)
    r = np.asarray(r_inverse, dtype=np.float64)
    t = np.asarray(t, dtype=np.float64)
    d0 = P[..., 0] - t[0]
    d1 = P[..., 1] - t[1]
    d2 = P[..., 2] - t[2]
    out = np.empty(P.shape, dtype=np.float64)
    for k in range(3):
        # "+ 0.0": np.dot's accumulators start at +0.0, so a sum of signed zeros is +0.0 in the reference (exact otherwise)
        out[..., k] = ((r[k, 0] * d0 + r[k, 1] * d1) + r[k, 2] * d2) + 0.0
    return out


def depth_to_world(raw, intr, r_inverse, t, mode=MODE_DEPTH, depth_scale=1.0, fB=0.0):
    """One frame of get_file_name's loop body (transfer/camera_to_world.py:160-166) without the text
    round trip (which is lossless): returns (cam (H*W,3), world (H*W,3)) fp64."""
    z = raw_to_z(raw, mode, depth_scale, fB)
    X, Y, Z = backproject(z, intr)
    cam = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    return cam, point_camera(cam, r_inverse, t)


def camera_centre(r_inverse, t):
    """Sensor origin for insertPointCloud: world position of p_cam = 0."""
    return point_camera(np.zeros((1, 3)), r_inverse, t)[0]


def valid_mask(raw, mode=MODE_DEPTH, depth_scale=1.0):
    """Pixels kept in compaction mode: raw decodes to a positive finite depth/disparity."""
    d = np.asarray(raw).astype(np.float64) * np.float64(depth_scale)
    return np.isfinite(d) & (d > 0)


# ----------------------------------------------------------------------------------------------
# Text formats
# ----------------------------------------------------------------------------------------------
def txt_lines_camera(X, Y, raw):
    """Bytes gentxtcord writes (camera_to_world.py:77-81): str(X),str(Y),str(Z)\\n with Z still the
    integer pixel value."""
    out = []
    Xf, Yf, Zr = X.ravel(), Y.ravel(), np.asarray(raw).ravel()
    for k in range(Xf.size):
        out.append(str(Xf[k]) + ',' + str(Yf[k]) + ',' + str(Zr[k]) + '\n')
    return "".join(out)


def txt_lines_world(world):
    """Bytes get_pointdata writes (camera_to_world.py:103-104)."""
    w = np.asarray(world, dtype=np.float64).reshape(-1, 3)
    return "".join(str(p[0]) + ',' + str(p[1]) + ',' + str(p[2]) + '\n' for p in w)


PLY_HEADER_XYZ = ("ply\n    format ascii 1.0\n    element vertex %d\n    property float x\n"
                  "    property float y\n    property float z\n    end_header\n    ")
PLY_TRAILER = "\n    "


def genply_text(x, y, z):
    """Exact text of genply (camera_to_world.py:112-134) / genply_RGB (pixel_to_camera.py:98-124)."""
    x = np.asarray(x, dtype=np.float64).ravel()
    y = np.asarray(y, dtype=np.float64).ravel()
    z = np.asarray(z, dtype=np.float64).ravel()
    body = "".join("%.4f %.4f %.4f \n" % (x[k], y[k], z[k]) for k in range(x.size))
    return (PLY_HEADER_XYZ % x.size) + body + PLY_TRAILER


# ----------------------------------------------------------------------------------------------
# Synthetic workloads -- shared by tests and bench so both see identical data
# ----------------------------------------------------------------------------------------------
def synth_pose(k, n, step=0.8):
    """Frame k of n: yaw = 0.2 sin(2 pi k/500) about camera-y, centre C = (2 sin(2 pi k/900), 0,
    step*(k - n/2)); returns Colmap world->cam (q scalar-last, t = -R C)."""
    yaw = 0.2 * np.sin(2 * np.pi * k / 500.0)
    c, s = np.cos(yaw), np.sin(yaw)
    Rwc = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], dtype=np.float64)  # cam -> world
    Rcw = Rwc.T
    C = np.array([2.0 * np.sin(2 * np.pi * k / 900.0), 0.0, step * (k - n / 2.0)])
    q = np.array([0.0, np.sin(-yaw / 2.0), 0.0, np.cos(-yaw / 2.0)])  # rotation about y by -yaw == Rcw
    t = -Rcw @ C
    return q, t


def synth_street_depth(W, H, intr, rng, noise=0.02):
    """Analytic 'street' range image in metres (float64): ground 1.65 m below the camera (+y down),
    facades at x = +-8 m, end wall at 80 m, sky above 12 m -> 0 (invalid)."""
    fx, fy, cx, cy = intr
    a = (np.arange(W) - cx) / fx
    b = (np.arange(H) - cy) / fy
    A, B = np.meshgrid(a, b)
    z = np.full((H, W), 80.0)
    with np.errstate(divide="ignore", invalid="ignore"):
        zg = np.where(B > 1e-9, 1.65 / B, np.inf)
        zw = np.where(np.abs(A) > 1e-9, 8.0 / np.abs(A), np.inf)
    z = np.minimum(z, np.minimum(zg, zw))
    sky = (-B * z) > 12.0
    z = z + rng.uniform(-noise, noise, size=z.shape)
    z[sky] = 0.0
    return np.clip(z, 0.0, 255.0)


def synth_depth_u16(W, H, intr, seed, kind="street"):
    """uint16 depth, metres = raw/256."""
    rng = np.random.default_rng(seed)
    if kind == "street":
        z = synth_street_depth(W, H, intr, rng)
    else:
        z = rng.uniform(1.0, 80.0, size=(H, W))
    return np.round(z * 256.0).astype(np.uint16)
