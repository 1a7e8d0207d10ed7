"""Host-side mirror of the reference's transfer/ scripts (same function names, arguments and files),
with every per-point loop executed by the CUDA kernels of libr3d_b200.so.

Reference                                   here
  pixel_to_camera.gentxtcord   :24-44        gentxtcord(filename, depth)            -> K1 (camera frame)
  camera_to_world.scipy_transfer :53-55      scipy_transfer(quat)                   -> r3d_pose_to_rt
  camera_to_world.point_camera :57-59        point_camera(p1, r_inverse, t)         -> r3d_pose_apply_points
  camera_to_world.get_pointdata :86-105      get_pointdata(p_path, q, t, x, y, z)   -> r3d_pose_apply_points
  camera_to_world.genply :112-134            genply / genply_RGB / genply_noRGB     (byte-identical ASCII PLY)
  camera_to_world.get_file_name :138-174     get_file_name(qt_path)                 -> one fused K1 batch

Module-level constants play the role of the constants the reference hard-codes in its sources.
"""
import os
import time

import numpy as np

from . import formats
from ._lib import MODE_DEPTH
from .runtime import default_context

# intrinsics hard-coded at camera_to_world.py:68-71 / pixel_to_camera.py:25-28
FX, FY, CX, CY = 600.391, 600.079, 320, 240
# paths hard-coded at camera_to_world.py:87,160,163,174
POINT_WORLD_PATH = './point_world/small_worldpoint_5_23_5.txt'
DEPTH_DIR = './depth/'
POINT_DIR = './point/'
PLY_PATH = './ply/small_035_p8.ply'
DEVICE = 0


def _intr(intr=None):
    return (FX, FY, CX, CY) if intr is None else tuple(intr)


def str_tofloat(data):
    """camera_to_world.py:28-30 (np.float was an alias of float)."""
    return np.array([float(v) for v in data])


def scipy_transfer(quat):
    """R^-1 of the scalar-last quaternion (normalised first), camera_to_world.py:53-55."""
    rt = default_context(DEVICE).pose_to_rt(np.asarray(quat, dtype=np.float64), np.zeros(3))
    return np.matrix(rt[0, :9].reshape(3, 3))


def point_camera(p1, r_inverse, t):
    """R^-1 (p - t) (camera_to_world.py:57-59).  A single point returns shape (3, 1) like the reference;
    an (n, 3) array returns (n, 3)."""
    p = np.asarray(p1, dtype=np.float64)
    rt = np.concatenate([np.asarray(r_inverse, dtype=np.float64).reshape(9), np.asarray(t, dtype=np.float64).reshape(3)])
    ctx = default_context(DEVICE)
    import ctypes as C  # noqa: F401
    from ._lib import check
    pts = np.ascontiguousarray(p.reshape(-1, 3))
    out = np.empty_like(pts)
    check(ctx.lib.r3d_pose_apply_points(ctx.handle, pts.ctypes.data, pts.shape[0], rt.ctypes.data, out.ctypes.data), ctx.handle)
    return out.reshape(3, 1) if p.ndim == 1 else out


def camera_points(depth, intr=None, mode=MODE_DEPTH, depth_scale=1.0, fB=0.0):
    """Back-projection of one frame in fp64: (H*W, 3) camera-frame points, row-major pixel order."""
    ctx = default_context(DEVICE)
    xyz, _ = ctx.backproject(np.ascontiguousarray(depth), _intr(intr), rt=None, mode=mode, depth_scale=depth_scale, fB=fB,
                             out_dtype=np.float64)
    return xyz


def gentxtcord(filename, depth, intr=None):
    """Depth image -> camera-frame points, written as `str(X),str(Y),str(Z)` lines (camera_to_world.py:67-83,
    pixel_to_camera.py:24-44).  Returns [xcord, ycord, zcord] like the pixel_to_camera variant (the
    camera_to_world variant returns None; its caller ignores the value)."""
    depth = np.ascontiguousarray(depth)
    xyz = camera_points(depth, intr)
    _write_xyz_txt(filename, xyz[:, 0], xyz[:, 1], xyz[:, 2], z_raw=depth)
    return [xyz[:, 0].tolist(), xyz[:, 1].tolist(), depth.ravel().tolist()]


def get_pointdata(p_path, q, t, xcord, ycord, zcord, point_world_path=None):
    """Re-read the camera-frame txt, move it to the world frame, append to the three coordinate lists and write the
    world txt (opened 'w' per call exactly like camera_to_world.py:86-105, so only the last frame survives)."""
    cam = formats.read_xyz_txt(p_path)
    r = scipy_transfer(q)
    world = point_camera(cam, r, t) if cam.shape[0] else cam
    xcord.extend(world[:, 0].tolist())
    ycord.extend(world[:, 1].tolist())
    zcord.extend(world[:, 2].tolist())
    _write_xyz_txt(point_world_path or POINT_WORLD_PATH, world[:, 0], world[:, 1], world[:, 2])


def _write_xyz_txt(path, x, y, z, z_raw=None):
    """`str(X),str(Y),str(Z)` lines formatted on the GPU (K6, r3d_format_txt_rows).  z_raw: integer samples printed as
    integers, as the reference does for camera-frame files (Z is still np.uint8 there); a float z_raw (not a reference
    case) goes through the host formatter."""
    if z_raw is not None:
        zr = np.asarray(z_raw).ravel()
        if zr.dtype.kind not in "iu":
            return formats.write_xyz_txt(path, x, y, z, z_raw=z_raw)
        rows = default_context(DEVICE).txt_rows(x, y, zr.astype(np.float64), z_is_integer=True)
    else:
        rows = default_context(DEVICE).txt_rows(x, y, z)
    with open(path, "wb") as f:
        f.write(rows)


def _write_ply(pc_file, x, y, z, rgb=None):
    """Header + GPU-formatted rows (K6, r3d_format_ply_rows) + trailer: the exact bytes of the reference's writers."""
    rows = default_context(DEVICE).ply_rows(x, y, z, rgb=rgb)
    hdr = (formats.PLY_HEADER_XYZ if rgb is None else formats.PLY_HEADER_RGB) % x.size
    with open(pc_file, "wb") as f:
        f.write(hdr.encode("ascii"))
        f.write(rows)
        f.write(formats.PLY_TRAILER.encode("ascii"))


def genply(gtxyz, pc_file, lenth_point):
    """ASCII PLY, byte-identical to camera_to_world.py:112-134."""
    x = np.asarray(gtxyz[0], dtype=np.float64)[:lenth_point]
    y = np.asarray(gtxyz[1], dtype=np.float64)[:lenth_point]
    z = np.asarray(gtxyz[2], dtype=np.float64)[:lenth_point]
    if not (x.size == y.size == z.size == lenth_point):
        raise ValueError("could not broadcast input array into shape (%d,)" % lenth_point)
    _write_ply(pc_file, x, y, z)
    print("Write into .ply file Done.")


def genply_RGB(gtxyz, pc_file):
    """pixel_to_camera.py:98-124 (despite the name: xyz only)."""
    genply(gtxyz, pc_file, len(gtxyz[0]))


def genply_noRGB(gtxyz, imgpath, pc_file):
    """pixel_to_camera.py:55-91 (despite the name: xyz + rgb + alpha 0)."""
    from PIL import Image
    t1 = time.time()
    img = np.array(Image.open(imgpath))
    n = img.shape[0] * img.shape[1]
    rgb = img[:, :, 0:3].reshape(n, 3)
    _write_ply(pc_file, np.asarray(gtxyz[0], dtype=np.float64)[:n], np.asarray(gtxyz[1], dtype=np.float64)[:n],
               np.asarray(gtxyz[2], dtype=np.float64)[:n], rgb=rgb)
    print("Write into .ply file Done.", time.time() - t1)


def sequence_to_world(depths, quats, trans, intr=None, mode=MODE_DEPTH, depth_scale=1.0, fB=0.0, t_scale=1.0,
                      out_dtype=np.float32, compact=False):
    """The fused in-memory path: (n, H, W) depth stack + poses -> (n*H*W, 3) world points in one kernel launch."""
    ctx = default_context(DEVICE)
    rt = ctx.pose_to_rt(quats, trans, t_scale=t_scale)
    return ctx.backproject(np.ascontiguousarray(depths), _intr(intr), rt=rt, mode=mode, depth_scale=depth_scale, fB=fB,
                           out_dtype=out_dtype, compact=compact)


def _rows_to_file(ctx, xyz_dev, n, path, kind, z_is_integer=False):
    """One small text file from device points (the per-frame side files of the reference)."""
    cap = n * (48 if kind == "ply" else 64) + 64
    buf = np.empty(cap, dtype=np.uint8)
    need = ctx.format_rows_into(xyz_dev, n, buf, kind, z_is_integer)
    if need > cap:
        buf = np.empty(need, dtype=np.uint8)
        need = ctx.format_rows_into(xyz_dev, n, buf, kind, z_is_integer)
    with open(path, "wb") as f:
        f.write(memoryview(buf)[:need])


def get_file_name(qt_path, intr=None, write_intermediate=True, ply_path=None, pose_format="comma", keep_points=False, max_frames=None,
                  quiet=False, frames_per_batch=64):
    """Sequence driver (camera_to_world.py:138-174): pose file -> per frame depth PNG -> world points -> one merged PLY,
    STREAMED: frame batches are decoded (IMREAD_GRAYSCALE semantics) into pinned memory on a worker thread while the batch
    before is on the GPU; a batch goes through ONE fused kernel launch per ring chunk, its float64 world points stay on the
    device and are formatted there (K6); only text comes back, and it is appended to the PLY by a writer thread while the
    next batch is formatted.  The PLY's vertex count is known up front (every pixel of every frame is a vertex), so the
    file is written front to back and host memory is bounded by two batches for any sequence length.
    write_intermediate keeps the reference's side files (./point/<name>.txt per frame, and the world txt, which the
    reference re-opens with 'w' for every frame so that only the last frame's survives).  Returns None like the reference,
    or (x, y, z) of every world point with keep_points=True."""
    from . import streaming
    poses = formats.read_pose_file(qt_path) if pose_format == "comma" else formats.read_colmap_images_txt(qt_path)
    say = (lambda *a: None) if quiet else print
    say('data start transfer')
    names = poses["names"] if max_frames is None else poses["names"][:max_frames]
    n = len(names)
    paths = [os.path.join(DEPTH_DIR, nm) for nm in names]
    ctx = default_context(DEVICE)
    total = sum(formats.frame_pixels(p) for p in paths)
    kept = []
    dec = streaming.BatchDecoder(paths, "gray", frames_per_batch)
    slots = streaming.TextSlots()
    with open(ply_path or PLY_PATH, "wb") as f:
        f.write((formats.PLY_HEADER_XYZ % total).encode("ascii"))
        writer = streaming.AsyncFileWriter(f)
        try:
            k = 0
            for stack, used in dec:
                t1 = time.time()
                j = k + used
                H, W = int(stack.shape[1]), int(stack.shape[2])
                npts = used * H * W
                rt = ctx.pose_to_rt(poses["q"][k:j], poses["t"][k:j])
                world = ctx.device_empty((npts, 3), np.float64)
                ctx.backproject(stack, _intr(intr), rt=rt, out=world)          # pinned stack -> staging ring -> device records
                slot, buf = slots.acquire(npts * 40 + 64)
                need = ctx.format_rows_into(world, npts, buf, "ply")
                if need > buf.size:
                    buf = slots.regrow(slot, need)
                    need = ctx.format_rows_into(world, npts, buf, "ply")
                writer.write(memoryview(buf)[:need], slots.releaser(slot))
                if keep_points:
                    kept.append(world.numpy())
                if write_intermediate:
                    cam = ctx.device_empty((npts, 3), np.float64)
                    ctx.backproject(stack, _intr(intr), rt=None, out=cam)
                    for i in range(used):
                        # Z is printed as the integer pixel value, as the reference does (it is still np.uint8 there)
                        _rows_to_file(ctx, cam[i * H * W:(i + 1) * H * W], H * W, os.path.join(POINT_DIR, names[k + i][0:-4] + '.txt'), "txt", True)
                    cam.free()
                    if j == n:
                        _rows_to_file(ctx, world[(used - 1) * H * W:], H * W, POINT_WORLD_PATH, "txt")
                world.free()
                t2 = time.time()
                say('##################')
                say("two epoch cost .", t2 - t1)
                say('the picture generation is: ', j)
                k = j
        finally:
            writer.close()
            slots.close()
            dec.close()
        f.write(formats.PLY_TRAILER.encode("ascii"))
    say("Write into .ply file Done.")
    if keep_points:
        w = np.concatenate(kept) if kept else np.zeros((0, 3))
        return w[:, 0].copy(), w[:, 1].copy(), w[:, 2].copy()
    return None
