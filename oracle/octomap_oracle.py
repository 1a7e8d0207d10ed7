"""ctypes front-end of oracle/octomap_oracle.c -- TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: restates the un-vendored, un-pinned `octomap` extension the reference
imports (octomap/txt_transfer_octomap.py:2, octomap/ply_transfer_octomap.py:2).  Exposes
the same call surface the reference scripts use (OcTree(res), updateNode(point, bool),
updateInnerOccupancy(), writeBinary(bytes)) plus insertPointCloud (upstream binding
signature) so parity tests read like the reference scripts.

Nothing under 3d_reconstruction_system_b200/ may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboctomap_oracle.so")
_lib = None


def build(force=False):
    """Compile the C restatement (gcc, seconds)."""
    src = os.path.join(_HERE, "octomap_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_SO)
    vp, sz, dbl, flt, i32 = C.c_void_p, C.c_size_t, C.c_double, C.c_float, C.c_int
    u16p = C.POINTER(C.c_uint16)
    fp = C.POINTER(C.c_float)
    L.oo_create.restype = vp
    L.oo_create.argtypes = [dbl]
    L.oo_destroy.argtypes = [vp]
    L.oo_clear.argtypes = [vp]
    L.oo_params.argtypes = [vp, fp]
    L.oo_size.restype = sz
    L.oo_size.argtypes = [vp]
    L.oo_coord_to_key.restype = i32
    L.oo_coord_to_key.argtypes = [vp, dbl, dbl, dbl, u16p]
    L.oo_key_to_coord.restype = dbl
    L.oo_key_to_coord.argtypes = [vp, C.c_uint16]
    L.oo_update_key_logodds.argtypes = [vp, u16p, flt]
    L.oo_update_key.argtypes = [vp, u16p, i32]
    L.oo_update_point.restype = i32
    L.oo_update_point.argtypes = [vp, dbl, dbl, dbl, i32]
    L.oo_update_point_logodds.restype = i32
    L.oo_update_point_logodds.argtypes = [vp, dbl, dbl, dbl, flt]
    L.oo_update_points.restype = sz
    L.oo_update_points.argtypes = [vp, vp, sz, i32]
    L.oo_update_points_f32.restype = sz
    L.oo_update_points_f32.argtypes = [vp, vp, sz, i32]
    L.oo_compute_ray_keys.restype = C.c_long
    L.oo_compute_ray_keys.argtypes = [vp, fp, fp, vp, C.c_long]
    L.oo_compute_update.argtypes = [vp, vp, sz, fp, dbl, C.POINTER(vp), C.POINTER(sz), C.POINTER(vp), C.POINTER(sz)]
    L.oo_free.argtypes = [vp]
    L.oo_insert_point_cloud_f32.argtypes = [vp, vp, sz, fp, dbl, i32]
    L.oo_insert_point_cloud.argtypes = [vp, vp, sz, vp, dbl, i32]
    L.oo_update_inner_occupancy.argtypes = [vp]
    L.oo_search.restype = i32
    L.oo_search.argtypes = [vp, u16p, fp]
    L.oo_leaves.restype = sz
    L.oo_leaves.argtypes = [vp, vp, vp, vp, sz]
    L.oo_to_max_likelihood.argtypes = [vp]
    L.oo_prune.argtypes = [vp]
    L.oo_write_binary_mem.restype = vp
    L.oo_write_binary_mem.argtypes = [vp, C.POINTER(sz)]
    L.oo_write_ot_mem.restype = vp
    L.oo_write_ot_mem.argtypes = [vp, C.POINTER(sz)]
    L.oo_write_binary.restype = i32
    L.oo_write_binary.argtypes = [vp, C.c_char_p]
    _lib = L
    return L


def _f3(v):
    return (C.c_float * 3)(float(np.float32(v[0])), float(np.float32(v[1])), float(np.float32(v[2])))


def unpack_keys(packed):
    """uint64 packed keys (kx | ky<<16 | kz<<32) -> (n,3) uint16."""
    packed = np.asarray(packed, dtype=np.uint64)
    out = np.empty((packed.size, 3), dtype=np.uint16)
    out[:, 0] = packed & np.uint64(0xFFFF)
    out[:, 1] = (packed >> np.uint64(16)) & np.uint64(0xFFFF)
    out[:, 2] = (packed >> np.uint64(32)) & np.uint64(0xFFFF)
    return out


def pack_keys(keys):
    keys = np.asarray(keys).astype(np.uint64).reshape(-1, 3)
    return keys[:, 0] | (keys[:, 1] << np.uint64(16)) | (keys[:, 2] << np.uint64(32))


class OcTree:
    """Oracle twin of octomap.OcTree (subset used by the reference + insertPointCloud)."""

    def __init__(self, resolution):
        self._L = lib()
        self._t = self._L.oo_create(float(resolution))
        self.resolution = float(resolution)

    def __del__(self):
        try:
            if self._t:
                self._L.oo_destroy(self._t)
                self._t = None
        except Exception:
            pass

    # --- reference call surface ---
    def updateNode(self, value, update, lazy_eval=False):
        v = np.asarray(value, dtype=np.float64)
        if isinstance(update, (bool, np.bool_)):
            return bool(self._L.oo_update_point(self._t, v[0], v[1], v[2], int(bool(update))))
        return bool(self._L.oo_update_point_logodds(self._t, v[0], v[1], v[2], float(update)))

    def updateNodes(self, points, occupied=True):
        p = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)
        return int(self._L.oo_update_points(self._t, p.ctypes.data, p.shape[0], int(bool(occupied))))

    def updateNodes_f32(self, points, occupied=True):
        p = np.ascontiguousarray(points, dtype=np.float32).reshape(-1, 3)
        return int(self._L.oo_update_points_f32(self._t, p.ctypes.data, p.shape[0], int(bool(occupied))))

    def insertPointCloud(self, pointcloud, origin, maxrange=-1.0, lazy_eval=False, discretize=False):
        p = np.ascontiguousarray(pointcloud, dtype=np.float64).reshape(-1, 3)
        o = np.ascontiguousarray(origin, dtype=np.float64)
        self._L.oo_insert_point_cloud(self._t, p.ctypes.data, p.shape[0], o.ctypes.data, float(maxrange), int(bool(discretize)))

    def updateInnerOccupancy(self):
        self._L.oo_update_inner_occupancy(self._t)

    def writeBinary(self, filename=None):
        if filename is None:
            return self.write_binary_bytes()
        if isinstance(filename, bytes):
            filename = filename.decode("utf-8")
        return bool(self._L.oo_write_binary(self._t, filename.encode("utf-8")))

    # --- extra probes used by the parity tests ---
    def write_binary_bytes(self):
        n = C.c_size_t(0)
        p = self._L.oo_write_binary_mem(self._t, C.byref(n))
        data = C.string_at(p, n.value)
        self._L.oo_free(p)
        return data

    def write_ot_bytes(self):
        """tree.write(): the full .ot serialisation (log-odds preserved); does not modify the tree."""
        n = C.c_size_t(0)
        p = self._L.oo_write_ot_mem(self._t, C.byref(n))
        data = C.string_at(p, n.value)
        self._L.oo_free(p)
        return data

    def size(self):
        return int(self._L.oo_size(self._t))

    def params(self):
        a = (C.c_float * 5)()
        self._L.oo_params(self._t, a)
        return dict(hit=a[0], miss=a[1], clamp_min=a[2], clamp_max=a[3], occ_thres=a[4])

    def coordToKey(self, p):
        k = (C.c_uint16 * 3)()
        ok = self._L.oo_coord_to_key(self._t, float(p[0]), float(p[1]), float(p[2]), k)
        return (int(k[0]), int(k[1]), int(k[2])) if ok else None

    def keyToCoord(self, k):
        return float(self._L.oo_key_to_coord(self._t, int(k)))

    def search(self, key):
        k = (C.c_uint16 * 3)(int(key[0]), int(key[1]), int(key[2]))
        v = C.c_float(0)
        ok = self._L.oo_search(self._t, k, C.byref(v))
        return float(v.value) if ok else None

    def computeRayKeys(self, origin, end, max_keys=200000):
        buf = np.zeros((max_keys, 3), dtype=np.uint16)
        n = self._L.oo_compute_ray_keys(self._t, _f3(origin), _f3(end), buf.ctypes.data, max_keys)
        if n < 0:
            return None
        assert n <= max_keys
        return buf[:n].copy()

    def computeUpdate(self, pointcloud, origin, maxrange=-1.0):
        """-> (free_packed_sorted, occ_packed_sorted) uint64 arrays, free already minus occupied."""
        p = np.ascontiguousarray(pointcloud, dtype=np.float32).reshape(-1, 3)
        fo, oo_ = C.c_void_p(), C.c_void_p()
        nf, no = C.c_size_t(0), C.c_size_t(0)
        self._L.oo_compute_update(self._t, p.ctypes.data, p.shape[0], _f3(origin), float(maxrange),
                                  C.byref(fo), C.byref(nf), C.byref(oo_), C.byref(no))
        fr = np.ctypeslib.as_array(C.cast(fo, C.POINTER(C.c_uint64)), shape=(max(nf.value, 1),))[:nf.value].copy()
        oc = np.ctypeslib.as_array(C.cast(oo_, C.POINTER(C.c_uint64)), shape=(max(no.value, 1),))[:no.value].copy()
        self._L.oo_free(fo)
        self._L.oo_free(oo_)
        return fr, oc

    def insertPointCloud_f32(self, pointcloud, origin, maxrange=-1.0, discretize=False):
        p = np.ascontiguousarray(pointcloud, dtype=np.float32).reshape(-1, 3)
        self._L.oo_insert_point_cloud_f32(self._t, p.ctypes.data, p.shape[0], _f3(origin), float(maxrange), int(bool(discretize)))

    def leaves(self):
        """-> (keys (n,3) uint16 min-corner, values float32, depths uint8) in pre-order."""
        n = int(self._L.oo_leaves(self._t, None, None, None, 0))
        keys = np.zeros((n, 3), dtype=np.uint16)
        vals = np.zeros(n, dtype=np.float32)
        depths = np.zeros(n, dtype=np.uint8)
        if n:
            self._L.oo_leaves(self._t, keys.ctypes.data, vals.ctypes.data, depths.ctypes.data, n)
        return keys, vals, depths

    def toMaxLikelihood(self):
        self._L.oo_to_max_likelihood(self._t)

    def prune(self):
        self._L.oo_prune(self._t)
