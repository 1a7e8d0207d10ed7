"""Thin object layer over the C ABI: one Context per GPU, buffers are numpy arrays (host) or anything with
data_ptr() / __cuda_array_interface__ (device, e.g. torch CUDA tensors used purely as memory holders)."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import (MODE_DEPTH, MODE_DISPARITY, OUT_F32, OUT_F64, check)

_NP_DTYPE = {np.dtype(np.uint8): _lib.U8, np.dtype(np.uint16): _lib.U16, np.dtype(np.float32): _lib.F32}


def _is_host(x):
    return isinstance(x, np.ndarray)


def _ptr(x):
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    if hasattr(x, "__cuda_array_interface__"):
        return x.__cuda_array_interface__["data"][0]
    if isinstance(x, int):
        return x
    raise TypeError("unsupported buffer type %r" % type(x))


def _depth_code(x):
    if isinstance(x, np.ndarray):
        dt = np.dtype(x.dtype)
        if dt not in _NP_DTYPE:
            raise TypeError("depth dtype must be uint8, uint16 or float32, got %s" % dt)
        return _NP_DTYPE[dt]
    name = str(getattr(x, "dtype", ""))
    for key, code in (("uint8", _lib.U8), ("uint16", _lib.U16), ("int16", _lib.U16), ("float32", _lib.F32)):
        if name.endswith(key):
            return code
    raise TypeError("cannot infer depth dtype from %r" % name)


class Context:
    """r3d_ctx: one per GPU; not thread-safe (the reference is single-threaded, synchronous)."""

    def __init__(self, device=0):
        self.lib = _lib.load()
        self.device = int(device)
        self._h = self.lib.r3d_create(self.device)
        if not self._h:
            msg = self.lib.r3d_last_error(None)
            raise _lib.R3DError("r3d_create(%d) failed: %s" % (device, msg.decode() if msg else "?"))

    def close(self):
        if getattr(self, "_h", None):
            self.lib.r3d_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- plumbing
    @property
    def handle(self):
        return self._h

    def stream(self):
        return self.lib.r3d_stream(self._h)

    def synchronize(self):
        check(self.lib.r3d_synchronize(self._h), self._h)

    def set_blocking(self, flag):
        check(self.lib.r3d_set_blocking(self._h, 1 if flag else 0), self._h)

    def launch_count(self):
        return int(self.lib.r3d_launch_count(self._h))

    def last_kernel_ms(self):
        return float(self.lib.r3d_last_kernel_ms(self._h))

    # ---- poses (scipy_transfer, transfer/camera_to_world.py:53-55)
    def pose_to_rt(self, quats, trans, t_scale=1.0):
        q = np.asarray(quats, dtype=np.float64).reshape(-1, 4)
        t = np.asarray(trans, dtype=np.float64).reshape(-1, 3)
        poses = np.ascontiguousarray(np.concatenate([q, t], axis=1))
        rt = np.empty((poses.shape[0], 12), dtype=np.float64)
        rc = self.lib.r3d_pose_to_rt(poses.ctypes.data, poses.shape[0], float(t_scale), rt.ctypes.data)
        if rc != 0:
            msg = self.lib.r3d_last_error(None).decode()
            if "zero norm" in msg:
                raise ValueError(msg)
            check(rc, None)
        return rt

    # ---- K1
    def backproject(self, depth, intr, rt=None, mode=MODE_DEPTH, depth_scale=1.0, fB=0.0, compact=False,
                    out=None, out_dtype=np.float32, shape=None, pitch=0, counts=None):
        """depth: (n, H, W) or (H, W) numpy array, or a device buffer with `shape=(n, H, W)` given.
        rt: (n, 12) float64 from pose_to_rt, or None for camera-frame points.
        Returns (xyz, counts): xyz (n*H*W, 3) [or the first sum(counts) rows in compact mode]."""
        if shape is None:
            shape = tuple(depth.shape)
        if len(shape) == 2:
            shape = (1,) + tuple(shape)
        n, H, W = (int(v) for v in shape)
        if _is_host(depth):
            depth = np.ascontiguousarray(depth)
        code = _depth_code(depth)
        out_np = np.dtype(out_dtype)
        ocode = OUT_F32 if out_np == np.dtype(np.float32) else OUT_F64
        if rt is not None and _is_host(rt):
            rt = np.ascontiguousarray(rt, dtype=np.float64).reshape(n, 12)
        ret_host = out is None
        if out is None:
            out = np.empty((n * H * W, 3), dtype=out_np)
        cnt = counts if counts is not None else np.zeros(max(n, 1), dtype=np.uint64)
        intr_a = (C.c_double * 4)(*[float(v) for v in intr])
        check(self.lib.r3d_backproject_rt(self._h, _ptr(depth), code, W, H, int(pitch), n, C.addressof(intr_a), _ptr(rt),
                                          int(mode), float(depth_scale), float(fB), 1 if compact else 0, ocode,
                                          _ptr(out), _ptr(cnt)), self._h)
        if ret_host and compact:
            out = out[: int(cnt[:n].sum())]
        return out, cnt[:n] if _is_host(cnt) else cnt

    def backproject_qt(self, depth, intr, quats, trans, **kw):
        """Same from quaternion (scalar-last) + translation poses: the r3d_backproject entry point."""
        rt = self.pose_to_rt(quats, trans)
        return self.backproject(depth, intr, rt=rt, **kw)

    def transform_points(self, xyz, T):
        """T . [x y z 1]^T for an (n,3) float64 cloud (other_tools/transfer_T_icp.py:10-12)."""
        p = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1, 3)
        Tm = np.ascontiguousarray(T, dtype=np.float64).reshape(16)
        out = np.empty_like(p)
        check(self.lib.r3d_transform_points(self._h, p.ctypes.data, p.shape[0], Tm.ctypes.data, out.ctypes.data), self._h)
        return out


_default = {}


def default_context(device=0):
    """Process-wide context per device (what the drop-in scripts use)."""
    if device not in _default:
        _default[device] = Context(device)
    return _default[device]
