"""Thin object layer over the C ABI: one Context per GPU, buffers are numpy arrays (host) or anything with
data_ptr() / __cuda_array_interface__ (device, e.g. torch CUDA tensors used purely as memory holders)."""
import ctypes as C
import weakref

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


class DeviceBuffer:
    """A block of device memory owned by a Context (r3d_device_alloc): shape / dtype bookkeeping plus data_ptr(), so it
    can be passed wherever the wrappers accept a device buffer.  Slicing along axis 0 gives a view."""

    def __init__(self, ctx, shape, dtype, _base=None, _ptr=None):
        self.ctx = ctx
        if isinstance(shape, (int, np.integer)):
            shape = (shape,)
        self.shape = tuple(int(v) for v in shape)
        self.dtype = np.dtype(dtype)
        self.nbytes = int(np.prod(self.shape, dtype=np.int64)) * self.dtype.itemsize
        self._base = _base
        if _base is None:
            self._p = ctx.lib.r3d_device_alloc(ctx.handle, self.nbytes)
            if not self._p:
                msg = ctx.lib.r3d_last_error(ctx.handle)
                raise _lib.R3DError("r3d_device_alloc(%d) failed: %s" % (self.nbytes, msg.decode() if msg else "?"))
        else:
            self._p = _ptr

    def data_ptr(self):
        return self._p

    def __getitem__(self, key):
        if not isinstance(key, slice):
            key = slice(int(key), int(key) + 1)
            squeeze = True
        else:
            squeeze = False
        a, b, step = key.indices(self.shape[0])
        if step != 1:
            raise ValueError("DeviceBuffer views are contiguous")
        b = max(a, b)
        row = self.nbytes // max(self.shape[0], 1)
        shape = (b - a,) + self.shape[1:]
        if squeeze:
            shape = self.shape[1:]
        return DeviceBuffer(self.ctx, shape, self.dtype, _base=self._base or self, _ptr=self._p + a * row)

    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        n = int(np.prod(self.shape, dtype=np.int64))
        shape = list(shape)
        if -1 in shape:
            i = shape.index(-1)
            rest = int(np.prod([v for v in shape if v != -1], dtype=np.int64))
            shape[i] = n // max(rest, 1)
        if int(np.prod(shape, dtype=np.int64)) != n:
            raise ValueError("cannot reshape %r to %r" % (self.shape, tuple(shape)))
        return DeviceBuffer(self.ctx, shape, self.dtype, _base=self._base or self, _ptr=self._p)

    def numpy(self):
        out = np.empty(self.shape, dtype=self.dtype)
        check(self.ctx.lib.r3d_memcpy(self.ctx.handle, out.ctypes.data, self._p, self.nbytes), self.ctx.handle)
        return out

    def copy_from(self, host):
        h = np.ascontiguousarray(host, dtype=self.dtype)
        if h.nbytes != self.nbytes:
            raise ValueError("size mismatch")
        check(self.ctx.lib.r3d_memcpy(self.ctx.handle, self._p, h.ctypes.data, self.nbytes), self.ctx.handle)
        return self

    def free(self):
        if self._base is None and getattr(self, "_p", None) and getattr(self.ctx, "_h", None):
            self.ctx.lib.r3d_device_free(self.ctx.handle, self._p)
        self._p = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    """r3d_ctx: one per GPU; not thread-safe (the reference is single-threaded, synchronous)."""

    def __init__(self, device=0):
        self.lib = _lib.load()
        self.device = int(device)
        self._children = weakref.WeakSet()     # objects whose native handles point into this context (trees)
        self._h = self.lib.r3d_create(self.device)
        if not self._h:
            msg = self.lib.r3d_last_error(None)
            raise _lib.R3DError("r3d_create(%d) failed: %s" % (device, msg.decode() if msg else "?"))

    def close(self):
        if getattr(self, "_h", None):
            for child in list(getattr(self, "_children", ())):      # native trees keep a pointer to the context: free them first
                child._release()
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

    def device_empty(self, shape, dtype):
        return DeviceBuffer(self, shape, dtype)

    def to_device(self, host):
        h = np.ascontiguousarray(host)
        return DeviceBuffer(self, h.shape, h.dtype).copy_from(h)

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
        if out is not None and str(getattr(out, "dtype", "")).endswith("float64"):
            out_np = np.dtype(np.float64)
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

    def ply_rows(self, x, y=None, z=None, rgb=None):
        """Vertex rows of the reference's ASCII PLY ("%.4f %.4f %.4f \\n", or with rgb "... r g b 0\\n") as bytes, formatted
        on the GPU (K6).  x alone may be an (n, 3) float64 array; or pass three equally long 1-D arrays."""
        if y is None:
            p = np.ascontiguousarray(x, dtype=np.float64).reshape(-1, 3)
            n, stride = p.shape[0], 3
            px, py, pz = p.ctypes.data, p.ctypes.data + 8, p.ctypes.data + 16
            keep = (p,)
        else:
            xa, ya, za = (np.ascontiguousarray(v, dtype=np.float64).ravel() for v in (x, y, z))
            if not (xa.size == ya.size == za.size):
                raise ValueError("coordinate arrays differ in length")
            n, stride = xa.size, 1
            px, py, pz = xa.ctypes.data, ya.ctypes.data, za.ctypes.data
            keep = (xa, ya, za)
        c = None
        if rgb is not None:
            c = np.ascontiguousarray(rgb, dtype=np.uint8).reshape(-1, 3)
            if c.shape[0] != n:
                raise ValueError("rgb must have one row per point")
        if n == 0:
            return b""
        need = C.c_size_t(0)
        # rows are at most 3*(1+17+5)+... bytes for |coordinates| < 1e17; size the buffer from a cheap bound, retry if short
        cap = n * (96 if c is None else 112)
        buf = np.empty(cap, dtype=np.uint8)
        check(self.lib.r3d_format_ply_rows(self._h, px, py, pz, stride, n, None if c is None else c.ctypes.data, buf.ctypes.data, cap,
                                           C.byref(need)), self._h)
        if need.value > cap:
            cap = need.value
            buf = np.empty(cap, dtype=np.uint8)
            check(self.lib.r3d_format_ply_rows(self._h, px, py, pz, stride, n, None if c is None else c.ctypes.data, buf.ctypes.data, cap,
                                               C.byref(need)), self._h)
        del keep
        return buf[: need.value].tobytes()

    def txt_rows(self, x, y=None, z=None, z_is_integer=False):
        """`str(X),str(Y),str(Z)\\n` lines of the reference's x,y,z txt files as bytes, formatted on the GPU (K6): every field
        is str(float64) (shortest round-trip repr); with z_is_integer Z is printed as an integer like gentxtcord does."""
        if y is None:
            p = np.ascontiguousarray(x, dtype=np.float64).reshape(-1, 3)
            n, stride = p.shape[0], 3
            px, py, pz = p.ctypes.data, p.ctypes.data + 8, p.ctypes.data + 16
            keep = (p,)
        else:
            xa, ya, za = (np.ascontiguousarray(v, dtype=np.float64).ravel() for v in (x, y, z))
            if not (xa.size == ya.size == za.size):
                raise ValueError("coordinate arrays differ in length")
            n, stride = xa.size, 1
            px, py, pz = xa.ctypes.data, ya.ctypes.data, za.ctypes.data
            keep = (xa, ya, za)
        if n == 0:
            return b""
        need = C.c_size_t(0)
        cap = n * 76
        buf = np.empty(cap, dtype=np.uint8)
        check(self.lib.r3d_format_txt_rows(self._h, px, py, pz, stride, n, 1 if z_is_integer else 0, buf.ctypes.data, cap, C.byref(need)), self._h)
        del keep
        return buf[: need.value].tobytes()

    def format_rows_into(self, xyz_dev, n, out, kind="ply", z_is_integer=False):
        """K6 from a DEVICE (n, 3) float64 buffer into a caller-provided host byte array (pinned memory from
        streaming.PinnedBuffer is the fast case): kind "ply" = genply's "%.4f %.4f %.4f \\n" rows, "txt" = the txt files'
        str(float64) rows.  Returns the number of bytes the rows take; when that exceeds out.size nothing was written
        (call again with a larger array)."""
        p = _ptr(xyz_dev)
        need = C.c_size_t(0)
        if kind == "ply":
            rc = self.lib.r3d_format_ply_rows(self._h, p, p + 8, p + 16, 3, int(n), None, out.ctypes.data, out.size, C.byref(need))
        else:
            rc = self.lib.r3d_format_txt_rows(self._h, p, p + 8, p + 16, 3, int(n), 1 if z_is_integer else 0, out.ctypes.data, out.size, C.byref(need))
        check(rc, self._h)
        return int(need.value)

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
