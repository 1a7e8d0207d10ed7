"""File formats either side of the hot path: pose files, depth PNGs, x,y,z txt, PLY, exactly as the
reference scripts read and write them."""
import os

import numpy as np

# ------------------------------------------------------------------------------------------------ poses


def read_pose_file(path):
    """The hand-preprocessed comma format get_file_name parses (transfer/camera_to_world.py:138-158):
    first line skipped; per line split(','): [0] id (ignored), [1:4] t, [4:8] q (fed as is to scipy's
    scalar-last from_quat), [8] depth PNG file name.  Returns dict(t (n,3), q (n,4), names list)."""
    t, q, names = [], [], []
    with open(path, "r") as f:
        f.readline()
        for line in f:
            if not line.strip():
                continue
            d = line.split(",")
            if len(d) < 9:
                raise ValueError("pose line needs >= 9 comma-separated fields: %r" % line)
            t.append([float(v) for v in d[1:4]])
            q.append([float(v) for v in d[4:8]])
            names.append(d[8].strip())
    return {"t": np.array(t, dtype=np.float64).reshape(-1, 3), "q": np.array(q, dtype=np.float64).reshape(-1, 4),
            "names": names}


def read_colmap_images_txt(path):
    """Raw Colmap images.txt (named by BASELINE.json north_star): `IMAGE_ID QW QX QY QZ TX TY TZ CAMERA_ID NAME`
    on every first line of a pair, `#` comments.  Quaternion reordered (w,x,y,z) -> scipy's (x,y,z,w)."""
    t, q, names = [], [], []
    with open(path, "r") as f:
        lines = [ln for ln in f if not ln.startswith("#")]
    i = 0
    while i < len(lines):
        parts = lines[i].split()
        if len(parts) >= 10:
            qw, qx, qy, qz = (float(v) for v in parts[1:5])
            q.append([qx, qy, qz, qw])
            t.append([float(v) for v in parts[5:8]])
            names.append(parts[9])
            i += 2   # the following line lists the 2-D points
        else:
            i += 1
    return {"t": np.array(t, dtype=np.float64).reshape(-1, 3), "q": np.array(q, dtype=np.float64).reshape(-1, 4),
            "names": names}


def write_pose_file(path, t, q, names):
    with open(path, "w") as f:
        f.write("id,tx,ty,tz,qx,qy,qz,qw,name,pad\n")
        for k, name in enumerate(names):
            f.write("%d,%r,%r,%r,%r,%r,%r,%r,%s,0\n" % ((k,) + tuple(float(v) for v in t[k]) + tuple(float(v) for v in q[k]) + (name,)))


# ------------------------------------------------------------------------------------------------ depth images


def png_info(path):
    """(W, H, channels as IMREAD_UNCHANGED shapes them (1 = no axis, 3, 4), bits per sample 8 | 16)."""
    import ctypes as C
    from . import _lib
    lib = _lib.load()
    w, h, c, d = C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0)
    rc = lib.r3d_png_info(os.fsencode(path), C.byref(w), C.byref(h), C.byref(c), C.byref(d))
    if rc != 0:
        msg = lib.r3d_last_error(None).decode("utf-8", "replace")
        if rc == -5:
            raise FileNotFoundError(msg)
        raise ValueError(msg)
    return w.value, h.value, c.value, d.value


def imread_batch(paths, mode="gray", channel=1, out=None, n_threads=0):
    """Decode a batch of equally sized PNG files on the native thread pool (r3d_png_decode_batch) into one (n, H, W) stack.
    mode: "gray" = cv.imread(p, IMREAD_GRAYSCALE); "channel" = cv.imread(p, IMREAD_UNCHANGED)[:, :, channel];
    "raw" = IMREAD_UNCHANGED, first channel of a multi-channel file.  out: optional preallocated (pinned) array."""
    import ctypes as C
    from . import _lib
    lib = _lib.load()
    paths = [os.fspath(p) for p in paths]
    n = len(paths)
    code = {"gray": _lib.PNG_GRAY8, "channel": _lib.PNG_CHANNEL, "raw": _lib.PNG_RAW}[mode]
    if n == 0:
        return np.zeros((0, 0, 0), dtype=np.uint8)
    w, h, _, depth = png_info(paths[0])
    dt = np.uint8 if (code == _lib.PNG_GRAY8 or depth == 8) else np.uint16
    if out is None:
        out = np.empty((n, h, w), dtype=dt)
    if out.shape != (n, h, w) or out.dtype != dt or not out.flags.c_contiguous:
        raise ValueError("out must be a C-contiguous %s array of shape %r" % (np.dtype(dt).name, (n, h, w)))
    arr = (C.c_char_p * n)(*[os.fsencode(p) for p in paths])
    rc = lib.r3d_png_decode_batch(arr, n, code, int(channel), out.ctypes.data, 0, out.dtype.itemsize, w, h, int(n_threads), None)
    if rc != 0:
        msg = lib.r3d_last_error(None).decode("utf-8", "replace")
        if "cannot open" in msg:
            raise FileNotFoundError(msg)
        if "no channel axis" in msg:
            raise IndexError(msg)
        raise ValueError(msg)
    return out


def read_frame_batch(paths, mode="gray", max_frames=256, alloc=None):
    """Longest prefix of `paths` (at most max_frames) that decodes to equally shaped frames of one dtype, as one (n, H, W)
    stack: PNG prefixes go through the thread-pool decoder in one call, anything else frame by frame.
    alloc(shape, dtype) -> array to decode into (e.g. a view of pinned memory); default: a fresh numpy array.
    Returns (stack, n_consumed)."""
    one = {"gray": imread_gray, "raw": imread_raw, "green": imread_unchanged_green}[mode]
    bmode, chan = {"gray": ("gray", 0), "raw": ("raw", 0), "green": ("channel", 1)}[mode]
    if not paths:
        return np.zeros((0, 0, 0), np.uint8), 0
    if _native_png(paths[0]):
        w, h, _, d = png_info(paths[0])
        n = 1
        while n < len(paths) and n < max_frames and _native_png(paths[n]):
            wn, hn, _, dn = png_info(paths[n])
            if (wn, hn, dn) != (w, h, d):
                break
            n += 1
        try:
            dst = None
            if alloc is not None:
                dst = alloc((n, h, w), np.uint8 if (bmode == "gray" or d == 8) else np.uint16)
            return imread_batch(paths[:n], bmode, channel=chan, out=dst), n
        except ValueError:
            pass        # a file of the prefix has a defect only the decode finds: frame by frame below (OpenCV has the last word)
    first = one(paths[0])
    batch = [first]
    while len(batch) < len(paths) and len(batch) < max_frames and not _native_png(paths[len(batch)]):
        img = one(paths[len(batch)])
        if img.shape != first.shape or img.dtype != first.dtype:
            break
        batch.append(img)
    if alloc is not None:
        dst = alloc((len(batch),) + first.shape, first.dtype)
        for i, img in enumerate(batch):
            dst[i] = img
        return dst, len(batch)
    return np.stack(batch), len(batch)


def frame_pixels(path):
    """W * H of a depth image without decoding it when it is a PNG the native decoder serves (header only)."""
    if _native_png(path):
        w, h, _, _ = png_info(path)
        return w * h
    img = imread_gray(path)
    return int(img.shape[0]) * int(img.shape[1])


def _native_png(path):
    """A PNG the native decoder serves (its header says so; interlaced files do not qualify)."""
    if not _is_png(path):
        return False
    try:
        png_info(path)
        return True
    except ValueError:
        return False


def _is_png(path):
    try:
        with open(path, "rb") as f:
            return f.read(8) == b"\x89PNG\r\n\x1a\n"
    except OSError:
        return False


def _native_or_cv2(path, mode, channel, cv2_read):
    """A PNG goes through the native decoder; a well-formed variant it does not serve (Adam7 interlacing, trailing bytes
    after the image data that libpng only warns about) is read by OpenCV -- the reference's reader -- instead, so every
    file cv.imread accepts is accepted here.  Other formats (the reference only uses PNG) always go through OpenCV."""
    if _is_png(path):
        try:
            return imread_batch([path], mode, channel=channel, n_threads=1)[0]
        except (ValueError, IndexError) as exc:
            try:
                img = cv2_read(path)
            except Exception:
                raise exc
            if img is None:
                raise exc
            return img
    img = cv2_read(path)
    if img is None:
        raise FileNotFoundError(path)
    return img


def _cv2_gray(path):
    import cv2
    return cv2.imread(path, cv2.IMREAD_GRAYSCALE)


def _cv2_green(path):
    import cv2
    img = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    return None if img is None else np.ascontiguousarray(img[:, :, 1])


def _cv2_raw(path):
    import cv2
    img = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    if img is None:
        return None
    return np.ascontiguousarray(img[:, :, 0] if img.ndim == 3 else img)


def imread_gray(path):
    """cv.imread(path, IMREAD_GRAYSCALE) (transfer/camera_to_world.py:160): always uint8 (16-bit -> >>8)."""
    return _native_or_cv2(path, "gray", 0, _cv2_gray)


def imread_unchanged_green(path):
    """cv.imread(path, IMREAD_UNCHANGED)[:, :, 1] (transfer/pixel_to_camera.py:133-134)."""
    return _native_or_cv2(path, "channel", 1, _cv2_green)


def imread_raw(path):
    """Full-precision single-channel depth / disparity (uint8 or uint16), for the metric pipelines."""
    return _native_or_cv2(path, "raw", 0, _cv2_raw)


# ------------------------------------------------------------------------------------------------ txt clouds


def _fmt_floats(vals):
    # str(np.float64) == repr(python float): shortest round-trip representation
    return [repr(v) for v in vals]


def write_xyz_txt(path, x, y, z, z_raw=None, mode="w"):
    """`str(X),str(Y),str(Z)\\n` per point (camera_to_world.py:79-81,103-104).  z_raw: integer pixel values printed
    as integers, as the reference does for camera-frame files (Z is still np.uint8 there)."""
    xs = _fmt_floats(np.asarray(x, dtype=np.float64).ravel().tolist())
    ys = _fmt_floats(np.asarray(y, dtype=np.float64).ravel().tolist())
    if z_raw is not None:
        zr = np.asarray(z_raw).ravel()
        if zr.dtype.kind in "iu":
            zs = [str(v) for v in zr.tolist()]
        else:
            zs = [str(v) for v in zr]   # numpy scalar str (float32 shortest repr)
    else:
        zs = _fmt_floats(np.asarray(z, dtype=np.float64).ravel().tolist())
    with open(path, mode) as f:
        f.write("".join([a + "," + b + "," + c + "\n" for a, b, c in zip(xs, ys, zs)]))


def _read_text_points(path, skip_lines, comma, max_points, n_threads=0):
    import ctypes as C
    from . import _lib
    lib = _lib.load()
    n = C.c_uint64(0)
    bp = os.fsencode(path)
    # one call: a row is at least 6 bytes ("0,0,0\n"), so size // 6 + 1 rows always fit (untouched pages cost nothing)
    try:
        bound = os.path.getsize(path) // 6 + 1
    except OSError:
        bound = 0
    if max_points:
        bound = min(bound, int(max_points))
    out = np.empty((bound, 3), dtype=np.float64)
    rc = lib.r3d_read_xyz_text(bp, int(skip_lines), 1 if comma else 0, int(max_points), out.ctypes.data if bound else None, bound, C.byref(n),
                               int(n_threads))
    if rc != 0:
        msg = lib.r3d_last_error(None).decode("utf-8", "replace")
        raise FileNotFoundError(msg) if "cannot open" in msg else ValueError(msg)
    if n.value > bound:          # cannot happen (see the bound); kept as the ABI's two-call protocol
        out = np.empty((n.value, 3), dtype=np.float64)
        rc = lib.r3d_read_xyz_text(bp, int(skip_lines), 1 if comma else 0, int(max_points), out.ctypes.data, n.value, C.byref(n), int(n_threads))
        if rc != 0:
            raise ValueError(lib.r3d_last_error(None).decode("utf-8", "replace"))
        return out
    return out[:n.value].copy()


def read_xyz_txt(path):
    """x,y,z per line, comma separated (octomap/txt_transfer_octomap.py:16-25; camera_to_world.py:92-98), parsed by the
    native thread pool (r3d_read_xyz_text)."""
    return _read_text_points(path, 0, True, 0)


# ------------------------------------------------------------------------------------------------ PLY

PLY_HEADER_XYZ = ("ply\n    format ascii 1.0\n    element vertex %d\n    property float x\n"
                  "    property float y\n    property float z\n    end_header\n    ")
PLY_HEADER_RGB = ("ply\n    format ascii 1.0\n    element vertex %d\n    property float x\n    property float y\n"
                  "    property float z\n    property uchar red\n    property uchar green\n    property uchar blue\n"
                  "    property uchar alpha\n    end_header\n    ")
PLY_TRAILER = "\n    "


def ply_ascii_text(x, y, z, rgb=None):
    """Exact text of genply (camera_to_world.py:112-134): indented header, '%.4f %.4f %.4f \\n' rows,
    trailing newline + 4 spaces.  rgb (n,3) -> the genply_noRGB variant (pixel_to_camera.py:55-91)."""
    x = np.asarray(x, dtype=np.float64).ravel()
    y = np.asarray(y, dtype=np.float64).ravel()
    z = np.asarray(z, dtype=np.float64).ravel()
    n = x.size
    chunks = []
    step = 1 << 16
    if rgb is None:
        for s in range(0, n, step):
            m = min(step, n - s)
            flat = np.stack([x[s:s + m], y[s:s + m], z[s:s + m]], axis=1).ravel().tolist()
            chunks.append(("%.4f %.4f %.4f \n" * m) % tuple(flat))
        return (PLY_HEADER_XYZ % n) + "".join(chunks) + PLY_TRAILER
    rgb = np.asarray(rgb).reshape(n, 3)
    for s in range(0, n, step):
        m = min(step, n - s)
        rows = []
        for k in range(s, s + m):
            rows.append("%.4f %.4f %.4f %d %d %d 0\n" % (x[k], y[k], z[k], int(rgb[k, 0]), int(rgb[k, 1]), int(rgb[k, 2])))
        chunks.append("".join(rows))
    return (PLY_HEADER_RGB % n) + "".join(chunks) + PLY_TRAILER


def write_ply_ascii(path, x, y, z, rgb=None):
    with open(path, "w") as f:
        f.write(ply_ascii_text(x, y, z, rgb))


def write_ply_binary(path, xyz_f32):
    """binary_little_endian PLY: the float32 records the back-projection kernel writes are the body as is."""
    xyz = np.ascontiguousarray(xyz_f32, dtype="<f4").reshape(-1, 3)
    hdr = ("ply\nformat binary_little_endian 1.0\nelement vertex %d\nproperty float x\nproperty float y\n"
           "property float z\nend_header\n" % xyz.shape[0])
    with open(path, "wb") as f:
        f.write(hdr.encode("ascii"))
        f.write(xyz.tobytes())


def read_ply_points(path, skip_lines=8, max_points=5400001):
    """txt_read of octomap/ply_transfer_octomap.py:16-40: skip exactly `skip_lines` lines, then whitespace-split rows
    (first three tokens used), stop after `max_points` points (the reference breaks at generation >= 5 400 000 after
    inserting that point).  Blank / indentation-only lines are skipped instead of raising (documented deviation)."""
    return _read_text_points(path, skip_lines, False, max_points)


def ensure_dir(path):
    d = os.path.dirname(path)
    if d and not os.path.isdir(d):
        os.makedirs(d, exist_ok=True)
