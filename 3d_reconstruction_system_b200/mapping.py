"""Host-side mirror of the reference's OctoMap scripts (octomap/txt_transfer_octomap.py,
octomap/ply_transfer_octomap.py = other_tools/ply_transfer_octomap.py) and of the north-star sequence mode
(depth frames + poses -> insertPointCloud per frame -> .bt).  Every per-point loop runs in the CUDA kernels of
libr3d_b200.so through the `octomap` drop-in class; this module only reads files and prints what the reference prints.
"""
import os

import numpy as np

from . import formats
from ._lib import MODE_DEPTH
from .octomap import OcTree
from .runtime import default_context

DEVICE = 0


def _progress(n):
    # the reference prints the counter before every 100 000th point (txt_transfer_octomap.py:26-27)
    for g in range(0, n, 100000):
        print('the generation: ', g)


def txt_read(file_path, tree):
    """octomap/txt_transfer_octomap.py:16-28: every `x,y,z` line -> tree.updateNode(point, True)."""
    pts = formats.read_xyz_txt(file_path)
    tree.updateNodes(pts, True)
    _progress(pts.shape[0])
    return pts.shape[0]


def ply_read(file_path, tree, skip_lines=8, max_points=5400001):
    """octomap/ply_transfer_octomap.py:16-40: skip 8 lines, whitespace-split rows -> updateNode(point, True), stop after
    the point with generation 5 400 000 (i.e. 5 400 001 points)."""
    pts = formats.read_ply_points(file_path, skip_lines=skip_lines, max_points=max_points)
    tree.updateNodes(pts, True)
    _progress(pts.shape[0])
    return pts.shape[0]


def cloud_file_to_bt(file_in, file_bt, resolution=0.1, kind="txt", device=None, **kw):
    """Whole script body (txt_...:31-36 / ply_...:43-48): OcTree(res), read + insert, updateInnerOccupancy, writeBinary."""
    tree = OcTree(resolution, device=DEVICE if device is None else device)
    n = txt_read(file_in, tree) if kind == "txt" else ply_read(file_in, tree, **kw)
    tree.updateInnerOccupancy()
    formats.ensure_dir(file_bt)
    ok = tree.writeBinary(bytes(file_bt, encoding='utf-8'))
    return tree, n, ok


def camera_centre(rt_row):
    """Sensor origin of a frame in the world: point_camera(0, R^-1, t) = R^-1 (0 - t) (camera_to_world.py:57-59),
    evaluated by the same kernel as every other point."""
    ctx = default_context(DEVICE)
    from ._lib import check
    p = np.zeros((1, 3), dtype=np.float64)
    out = np.empty_like(p)
    rt = np.ascontiguousarray(rt_row, dtype=np.float64).reshape(12)
    check(ctx.lib.r3d_pose_apply_points(ctx.handle, p.ctypes.data, 1, rt.ctypes.data, out.ctypes.data), ctx.handle)
    return out[0]


def camera_centres(rt):
    """Sensor origins of n frames at once: R^-1 (0 - t) per pose-table row, in the kernel's operation order (every product and
    sum rounded separately, left to right, + 0.0 last: r3d_math.cuh::pose_apply / pose_canon), so the values equal what
    camera_centre() gets from the GPU.  Per-frame host arithmetic, like r3d_pose_to_rt."""
    rt = np.ascontiguousarray(rt, dtype=np.float64).reshape(-1, 12)
    d0, d1, d2 = 0.0 - rt[:, 9], 0.0 - rt[:, 10], 0.0 - rt[:, 11]
    out = np.empty((rt.shape[0], 3), dtype=np.float64)
    for k in range(3):
        out[:, k] = ((rt[:, 3 * k] * d0 + rt[:, 3 * k + 1] * d1) + rt[:, 3 * k + 2] * d2) + 0.0
    return out


def sequence_to_octree(depths, quats, trans, intr, resolution=0.1, maxrange=80.0, mode=MODE_DEPTH, depth_scale=1.0, fB=0.0,
                       t_scale=1.0, tree=None, drop_invalid=True, frames_per_batch=64):
    """North-star mode (BASELINE.json configs 3-5): every frame is one scan.  Frames are back-projected to float32 world
    points by the fused kernel in batches that stay on the GPU, then inserted in frame order by ONE pipelined library call
    per batch: insertPointCloud(points, origin = camera centre, maxrange) per frame.

    drop_invalid (default): pixels whose decoded depth / disparity is not positive (sky, holes -- routine in disparity maps)
    are compacted away on the device (K1's compaction mode, per-frame counts read back once per batch).  The reference has
    no validity filter; with drop_invalid=False such pixels stay what its back-projection makes of them -- points AT the
    camera centre, which mark the sensor's own voxel occupied in every frame and leave an occupied trail along the
    trajectory in the .bt."""
    ctx = default_context(DEVICE)
    depths = np.ascontiguousarray(depths)
    n, H, W = depths.shape
    rt = ctx.pose_to_rt(quats, trans, t_scale=t_scale)
    origins = camera_centres(rt)
    if tree is None:
        tree = OcTree(resolution, ctx=ctx)
    for a in range(0, n, frames_per_batch):
        b = min(n, a + frames_per_batch)
        # the batch's world points never leave the GPU between the two kernels
        xyz = ctx.device_empty(((b - a) * H * W, 3), np.float32)
        if drop_invalid:
            _, cnt = ctx.backproject(depths[a:b], intr, rt=rt[a:b], mode=mode, depth_scale=depth_scale, fB=fB, out=xyz, compact=True)
            total = int(cnt.sum())
            tree.insertPointClouds(xyz[:total] if total else xyz[:0], origins[a:b], maxrange=maxrange, counts=cnt)
        else:
            ctx.backproject(depths[a:b], intr, rt=rt[a:b], mode=mode, depth_scale=depth_scale, fB=fB, out=xyz)
            tree.insertPointClouds(xyz, origins[a:b], maxrange=maxrange)
        xyz.free()
    return tree


def pose_sequence_to_bt(qt_path, file_bt, intr, depth_dir='./depth/', resolution=0.1, maxrange=80.0, pose_format="comma",
                        raw_depth=False, **kw):
    """Pose file + depth PNGs -> .bt, the file-level form of sequence_to_octree, streamed (decode overlaps the GPU work)."""
    from . import streaming
    poses = formats.read_pose_file(qt_path) if pose_format == "comma" else formats.read_colmap_images_txt(qt_path)
    tree = None
    # batch k + 1 is decoded into pinned memory on a worker thread while batch k is back-projected and ray-cast
    dec = streaming.BatchDecoder([os.path.join(depth_dir, nm) for nm in poses["names"]], "raw" if raw_depth else "gray", 64)
    try:
        k = 0
        for stack, used in dec:
            tree = sequence_to_octree(stack, poses["q"][k:k + used], poses["t"][k:k + used], intr, resolution=resolution, maxrange=maxrange,
                                      tree=tree, **kw)
            k += used
    finally:
        dec.close()
    if tree is None:
        tree = OcTree(resolution, device=DEVICE)
    tree.updateInnerOccupancy()
    formats.ensure_dir(file_bt)
    tree.writeBinary(bytes(file_bt, encoding='utf-8'))
    return tree
