"""Drop-in for the subset of the `octomap` Python module the reference scripts use
(octomap/txt_transfer_octomap.py:2,25,33-36; octomap/ply_transfer_octomap.py:2,33,45-48):

    tree = octomap.OcTree(0.1)
    tree.updateNode(point, True)          # per point, as the reference does
    tree.updateInnerOccupancy()
    tree.writeBinary(bytes(path, 'utf-8'))

plus insertPointCloud(points, origin, maxrange=-1., lazy_eval=False, discretize=False) with the upstream binding's
signature (named by BASELINE.json north_star).  Everything executes on the GPU through libr3d_b200.so.

Per-point updateNode calls are queued on the host and flushed to the GPU in batches, in call order; a batch of
updates with the same log-odds increment is order-independent per voxel, so runs of equal increments are applied
as one kernel launch and runs are applied in sequence -- the result equals the one-by-one loop.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check
from .runtime import _ptr, default_context

_FLUSH_POINTS = 1 << 20


class OcTree:
    def __init__(self, resolution, device=0, ctx=None):
        self._ctx = ctx if ctx is not None else default_context(device)
        self._lib = self._ctx.lib
        h = C.c_void_p()
        check(self._lib.r3d_tree_create(self._ctx.handle, float(resolution), C.byref(h)), self._ctx.handle)
        self._h = h
        self._ctx._children.add(self)
        self._res = float(resolution)
        self._pend_pts = []      # queued updateNode points (float64 triples)
        self._pend_upd = []      # their log-odds increments (float32)
        p = (C.c_float * 5)()
        check(self._lib.r3d_tree_params(self._h, p), self._ctx.handle)
        self._hit, self._miss, self._cmin, self._cmax, self._thres = (np.float32(v) for v in p)
        self.n_dropped = 0       # out-of-range points silently ignored, as upstream does

    def _release(self):
        if getattr(self, "_h", None):
            self._lib.r3d_tree_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    # ------------------------------------------------------------------ reference call surface
    def updateNode(self, value, update, lazy_eval=False):
        """updateNode(point, bool) / updateNode(point, float): dispatch on the Python type of `update` exactly like
        the upstream binding; the point is cast to float32 and only elements 0..2 are used."""
        if isinstance(update, (bool, np.bool_)):
            upd = self._hit if update else self._miss
        else:
            upd = np.float32(update)
        self._pend_pts.append((float(value[0]), float(value[1]), float(value[2])))
        self._pend_upd.append(upd)
        if len(self._pend_pts) >= _FLUSH_POINTS:
            self._flush()

    def updateNodes(self, points, occupied=True):
        """Batched form of the reference's per-point loop: n points, one increment."""
        self._flush()
        dropped = C.c_uint64(0)
        if isinstance(points, np.ndarray) and points.dtype == np.float32:
            p = np.ascontiguousarray(points).reshape(-1, 3)
            check(self._lib.r3d_tree_update_points(self._h, p.ctypes.data, p.shape[0], 1 if occupied else 0, C.byref(dropped)), self._ctx.handle)
        elif isinstance(points, np.ndarray) or isinstance(points, (list, tuple)):
            p = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)
            check(self._lib.r3d_tree_update_points_f64(self._h, p.ctypes.data, p.shape[0], 1 if occupied else 0, C.byref(dropped)), self._ctx.handle)
        else:   # device float32 buffer (n, 3)
            n = int(points.shape[0])
            check(self._lib.r3d_tree_update_points(self._h, _ptr(points), n, 1 if occupied else 0, C.byref(dropped)), self._ctx.handle)
        self.n_dropped += int(dropped.value)

    def insertPointCloud(self, pointcloud, origin, maxrange=-1.0, lazy_eval=False, discretize=False):
        self._flush()
        o = np.asarray(origin, dtype=np.float64).astype(np.float32)
        oa = (C.c_float * 3)(float(o[0]), float(o[1]), float(o[2]))
        if isinstance(pointcloud, np.ndarray) or isinstance(pointcloud, (list, tuple)):
            p = np.ascontiguousarray(np.asarray(pointcloud, dtype=np.float64).astype(np.float32)).reshape(-1, 3)
            ptr, n = p.ctypes.data, p.shape[0]
        else:   # device float32 (n, 3)
            ptr, n = _ptr(pointcloud), int(pointcloud.shape[0])
        check(self._lib.r3d_tree_insert_scan(self._h, ptr, n, oa, float(maxrange), 1 if discretize else 0), self._ctx.handle)

    def insertPointClouds(self, pointclouds, origins, maxrange=-1.0, discretize=False, counts=None):
        """A sequence of insertPointCloud calls in one library call.  pointclouds: (S, N, 3) float32 host array or device
        buffer (scans back to back), or a flat (sum N_s, 3) buffer with `counts` giving N_s; origins: (S, 3)."""
        self._flush()
        o = np.ascontiguousarray(np.asarray(origins, dtype=np.float64).astype(np.float32)).reshape(-1, 3)
        S = o.shape[0]
        if isinstance(pointclouds, np.ndarray):
            p = np.ascontiguousarray(pointclouds, dtype=np.float32)
            ptr = p.ctypes.data
            total = p.size // 3
        else:
            p = pointclouds
            ptr = _ptr(p)
            total = int(np.prod(p.shape)) // 3
        if counts is None:
            if S == 0 or total % S:
                raise ValueError("pointclouds do not divide into %d equal scans; pass counts" % S)
            cnt = np.full(S, total // S, dtype=np.uint64)
        else:
            cnt = np.ascontiguousarray(counts, dtype=np.uint64)
            if cnt.size != S or int(cnt.sum()) != total:
                raise ValueError("counts do not match the point buffer")
        check(self._lib.r3d_tree_insert_scans(self._h, ptr, cnt.ctypes.data, o.ctypes.data, S, float(maxrange), 1 if discretize else 0),
              self._ctx.handle)

    def lastScanStats(self):
        """dict(rays, steps, records, bricks) of the most recent insertPointCloud / computeScanDelta."""
        a = (C.c_uint64 * 4)()
        check(self._lib.r3d_tree_last_scan_stats(self._h, a), self._ctx.handle)
        return {"rays": int(a[0]), "steps": int(a[1]), "records": int(a[2]), "bricks": int(a[3])}

    def pipelineStats(self):
        """dict(wait_ms, work_ms, max_turnaround_ms, scans): the host's side of the last pipelined insertPointClouds batch."""
        a = (C.c_uint64 * 4)()
        check(self._lib.r3d_tree_pipeline_stats(self._h, a), self._ctx.handle)
        return {"wait_ms": a[0] / 1e6, "work_ms": a[1] / 1e6, "max_turnaround_ms": a[2] / 1e6, "scans": int(a[3])}

    def growthStats(self):
        """dict(pool_regrowths, table_regrowths, pool_capacity_bricks, table_capacity) since the tree was created."""
        a = (C.c_uint64 * 4)()
        check(self._lib.r3d_tree_growth_stats(self._h, a), self._ctx.handle)
        return {"pool_regrowths": int(a[0]), "table_regrowths": int(a[1]), "pool_capacity_bricks": int(a[2]), "table_capacity": int(a[3])}

    def updateInnerOccupancy(self):
        self._flush()
        check(self._lib.r3d_tree_update_inner_occupancy(self._h), self._ctx.handle)

    def _serialise(self, fn, per_brick):
        """bytes of a *_mem writer in ONE serialisation: the buffer is sized for the worst case (untouched pages cost
        nothing) -- .bt: 73 inner nodes of 2 bytes per brick plus the levels above; .ot: 585 nodes of 5 bytes."""
        n = C.c_size_t(0)
        cap = min((1 << 16) + per_brick * int(self.numBricks()), 1 << 30)     # (a larger file takes a second pass)
        while True:
            buf = np.empty(cap, dtype=np.uint8)
            check(fn(self._h, buf.ctypes.data, cap, C.byref(n)), self._ctx.handle)
            if n.value <= cap:
                return buf[: n.value].tobytes()
            cap = n.value + n.value // 8

    def writeBinary(self, filename=None):
        """writeBinary(bytes path) -> bool; with no argument returns the .bt bytes (the binding's overload)."""
        self._flush()
        if filename is None:
            return self._serialise(self._lib.r3d_tree_write_bt_mem, 170)
        if isinstance(filename, str):
            filename = filename.encode("utf-8")
        rc = self._lib.r3d_tree_write_bt(self._h, filename)
        if rc == -5:
            return False
        check(rc, self._ctx.handle)
        return True

    def readBinary(self, source):
        """readBinary(bytes path | str path) -> bool, or readBinary(file content as bytes starting with '# Octomap'):
        replaces the tree (and its resolution) with the .bt file's content."""
        self._pend_pts, self._pend_upd = [], []
        if isinstance(source, (bytes, bytearray)) and bytes(source[:9]) == b"# Octomap":
            buf = np.frombuffer(bytes(source), dtype=np.uint8)
            check(self._lib.r3d_tree_read_bt_mem(self._h, buf.ctypes.data, buf.size), self._ctx.handle)
        else:
            if isinstance(source, str):
                source = source.encode("utf-8")
            rc = self._lib.r3d_tree_read_bt(self._h, bytes(source))
            if rc == -5 and b"cannot open" in self._lib.r3d_last_error(self._ctx.handle):
                return False
            check(rc, self._ctx.handle)
        p = (C.c_float * 5)()
        check(self._lib.r3d_tree_params(self._h, p), self._ctx.handle)
        self._res = float(self.getResolutionFromLibrary())
        return True

    def write(self, filename=None):
        """write(bytes path) -> bool: the .ot format (every node's log-odds kept); with no argument returns the bytes."""
        self._flush()
        if filename is None:
            return self._serialise(self._lib.r3d_tree_write_ot_mem, 3400)
        if isinstance(filename, str):
            filename = filename.encode("utf-8")
        rc = self._lib.r3d_tree_write_ot(self._h, filename)
        if rc == -5:
            return False
        check(rc, self._ctx.handle)
        return True

    def read(self, source):
        """read(bytes path | str path), or read(.ot file content starting with '# Octomap'): replaces the tree."""
        self._pend_pts, self._pend_upd = [], []
        if isinstance(source, (bytes, bytearray)) and bytes(source[:9]) == b"# Octomap":
            buf = np.frombuffer(bytes(source), dtype=np.uint8)
            check(self._lib.r3d_tree_read_ot_mem(self._h, buf.ctypes.data, buf.size), self._ctx.handle)
        else:
            if isinstance(source, str):
                source = source.encode("utf-8")
            rc = self._lib.r3d_tree_read_ot(self._h, bytes(source))
            if rc == -5 and b"cannot open" in self._lib.r3d_last_error(self._ctx.handle):
                return False
            check(rc, self._ctx.handle)
        self._res = float(self.getResolutionFromLibrary())
        return True

    def getResolutionFromLibrary(self):
        r = C.c_double(0)
        check(self._lib.r3d_tree_resolution(self._h, C.byref(r)), self._ctx.handle)
        return r.value

    # ------------------------------------------------------------------ further binding methods
    def getResolution(self):
        return self._res

    def size(self):
        self._flush()
        n = C.c_uint64(0)
        check(self._lib.r3d_tree_size(self._h, C.byref(n)), self._ctx.handle)
        return int(n.value)

    def clear(self):
        self._pend_pts, self._pend_upd = [], []
        check(self._lib.r3d_tree_clear(self._h), self._ctx.handle)

    def reserve(self, n_bricks):
        """Capacity hint: room for n_bricks bricks (8x8x8 voxels each) without regrowing the device pool."""
        check(self._lib.r3d_tree_reserve(self._h, int(n_bricks)), self._ctx.handle)

    def toMaxLikelihood(self):
        self._flush()
        check(self._lib.r3d_tree_to_max_likelihood(self._h), self._ctx.handle)

    def coordToKey(self, point):
        p = np.asarray(point, dtype=np.float64).astype(np.float32).reshape(1, 3)
        k = np.zeros((1, 3), np.uint16)
        v = np.zeros(1, np.uint8)
        check(self._lib.r3d_coord_to_key(self._h, p.ctypes.data, 1, k.ctypes.data, v.ctypes.data), self._ctx.handle)
        return (int(k[0, 0]), int(k[0, 1]), int(k[0, 2])) if v[0] else None

    def coordsToKeys(self, points):
        p = np.ascontiguousarray(np.asarray(points, dtype=np.float64).astype(np.float32)).reshape(-1, 3)
        k = np.zeros((p.shape[0], 3), np.uint16)
        v = np.zeros(p.shape[0], np.uint8)
        check(self._lib.r3d_coord_to_key(self._h, p.ctypes.data, p.shape[0], k.ctypes.data, v.ctypes.data), self._ctx.handle)
        return k, v.astype(bool)

    def search(self, key):
        """Log-odds of the depth-16 leaf at `key` (3 x uint16), or None when unknown."""
        vals, found = self.searchKeys(np.asarray(key, dtype=np.uint16).reshape(1, 3))
        return float(vals[0]) if found[0] else None

    def searchKeys(self, keys):
        self._flush()
        k = np.ascontiguousarray(keys, dtype=np.uint16).reshape(-1, 3)
        vals = np.zeros(k.shape[0], np.float32)
        found = np.zeros(k.shape[0], np.uint8)
        check(self._lib.r3d_tree_search(self._h, k.ctypes.data, k.shape[0], vals.ctypes.data, found.ctypes.data), self._ctx.handle)
        return vals, found.astype(bool)

    def isNodeOccupied(self, value):
        return value is not None and np.float32(value) >= self._thres

    def numVoxels(self):
        self._flush()
        n = C.c_uint64(0)
        check(self._lib.r3d_tree_num_voxels(self._h, C.byref(n)), self._ctx.handle)
        return int(n.value)

    def voxels(self):
        """Every depth-16 leaf ever updated: (keys (n,3) uint16, log-odds float32), sorted by packed key."""
        n = self.numVoxels()
        keys = np.zeros((max(n, 1), 3), np.uint16)
        vals = np.zeros(max(n, 1), np.float32)
        m = C.c_uint64(0)
        check(self._lib.r3d_tree_export_voxels(self._h, keys.ctypes.data, vals.ctypes.data, n, C.byref(m)), self._ctx.handle)
        keys, vals = keys[:n], vals[:n]
        packed = keys[:, 0].astype(np.uint64) | (keys[:, 1].astype(np.uint64) << np.uint64(16)) | (keys[:, 2].astype(np.uint64) << np.uint64(32))
        order = np.argsort(packed, kind="stable")
        return keys[order], vals[order]

    # ------------------------------------------------------------------ scan deltas (multi-GPU merge)
    def computeScanDelta(self, pointcloud, origin, maxrange=-1.0, discretize=False):
        """Ray-cast one scan WITHOUT touching the tree; returns the delta as a (n, 136) uint8 array of brick records."""
        self._flush()
        o = np.asarray(origin, dtype=np.float64).astype(np.float32)
        oa = (C.c_float * 3)(float(o[0]), float(o[1]), float(o[2]))
        if isinstance(pointcloud, np.ndarray) or isinstance(pointcloud, (list, tuple)):
            p = np.ascontiguousarray(np.asarray(pointcloud, dtype=np.float64).astype(np.float32)).reshape(-1, 3)
            ptr, n = p.ctypes.data, p.shape[0]
        else:
            ptr, n = _ptr(pointcloud), int(pointcloud.shape[0])
        cnt = C.c_uint64(0)
        check(self._lib.r3d_scan_delta_compute(self._h, ptr, n, oa, float(maxrange), 1 if discretize else 0, C.byref(cnt)), self._ctx.handle)
        rec = np.zeros((max(cnt.value, 1), _lib.DELTA_RECORD_BYTES), np.uint8)
        check(self._lib.r3d_scan_delta_export(self._h, rec.ctypes.data, cnt.value, C.byref(cnt)), self._ctx.handle)
        return rec[: cnt.value]

    def computeScanDeltaOnDevice(self, pointcloud, origin, maxrange=-1.0, discretize=False):
        """Ray-cast one scan into the tree's device-side delta buffer (no host copy); returns the record count.
        Follow with scanDeltaInto(device_buffer, n)."""
        self._flush()
        o = np.asarray(origin, dtype=np.float64).astype(np.float32)
        oa = (C.c_float * 3)(float(o[0]), float(o[1]), float(o[2]))
        if isinstance(pointcloud, np.ndarray) or isinstance(pointcloud, (list, tuple)):
            p = np.ascontiguousarray(np.asarray(pointcloud, dtype=np.float64).astype(np.float32)).reshape(-1, 3)
            ptr, n = p.ctypes.data, p.shape[0]
        else:
            ptr, n = _ptr(pointcloud), int(pointcloud.shape[0])
        cnt = C.c_uint64(0)
        check(self._lib.r3d_scan_delta_compute(self._h, ptr, n, oa, float(maxrange), 1 if discretize else 0, C.byref(cnt)), self._ctx.handle)
        return int(cnt.value)

    def computeScanDeltasInto(self, pointclouds, counts, origins, maxrange, out_ptr, capacity_records, discretize=False):
        """Ray-cast len(origins) scans (points back to back in one float32 device / host buffer, `counts` points each)
        into the record buffer at out_ptr.  Returns per-scan record counts; raises MemoryError when the buffer is too
        small (the counts so far are in the exception's args[1])."""
        self._flush()
        o = np.ascontiguousarray(np.asarray(origins, dtype=np.float64).astype(np.float32)).reshape(-1, 3)
        cnt = np.ascontiguousarray(counts, dtype=np.uint64)
        rec = np.zeros(o.shape[0], dtype=np.uint64)
        ptr = pointclouds.ctypes.data if isinstance(pointclouds, np.ndarray) else _ptr(pointclouds)
        rc = self._lib.r3d_scan_deltas_compute(self._h, ptr, cnt.ctypes.data, o.ctypes.data, o.shape[0], float(maxrange), 1 if discretize else 0,
                                               _ptr(out_ptr), int(capacity_records), rec.ctypes.data)
        if rc == -3:
            raise MemoryError(self._lib.r3d_last_error(self._ctx.handle).decode(), rec)
        check(rc, self._ctx.handle)
        return rec

    def applyDeltasOwned(self, records_ptr, counts, part=0, nparts=1):
        """Apply consecutive deltas (device records back to back, counts[s] each) in order, restricted to owned bricks."""
        self._flush()
        cnt = np.ascontiguousarray(counts, dtype=np.uint64)
        check(self._lib.r3d_tree_apply_deltas_owned(self._h, _ptr(records_ptr), cnt.ctypes.data, cnt.size, int(part), int(nparts)), self._ctx.handle)

    def deferDeltasOwned(self, records_ptr, counts, part=0, nparts=1):
        """applyDeltasOwned, but only noted: everything noted is applied in ONE sorted, scan-ordered pass by the next
        computeScanDeltasInto (or whatever touches the map first).  The records must stay valid until then."""
        self._flush()
        cnt = np.ascontiguousarray(counts, dtype=np.uint64)
        check(self._lib.r3d_tree_defer_deltas_owned(self._h, _ptr(records_ptr), cnt.ctypes.data, cnt.size, int(part), int(nparts)), self._ctx.handle)

    def flushDeferred(self):
        check(self._lib.r3d_tree_flush_deferred(self._h), self._ctx.handle)

    def scanDeltaInto(self, buf, capacity_records):
        """Export the last computed delta into a caller buffer (host array or device tensor); returns the record count."""
        cnt = C.c_uint64(0)
        check(self._lib.r3d_scan_delta_export(self._h, _ptr(buf), int(capacity_records), C.byref(cnt)), self._ctx.handle)
        return int(cnt.value)

    def applyDelta(self, records, n_records=None):
        self._flush()
        if isinstance(records, np.ndarray):
            records = np.ascontiguousarray(records, dtype=np.uint8)
            n = records.size // _lib.DELTA_RECORD_BYTES if n_records is None else int(n_records)
            ptr = records.ctypes.data
        else:
            ptr, n = _ptr(records), int(n_records)
        check(self._lib.r3d_tree_apply_delta(self._h, ptr, n), self._ctx.handle)

    def applyDeltaOwned(self, records, n_records, part, nparts):
        """applyDelta restricted to the bricks owned by partition `part` of `nparts` (multi-GPU partitioned map)."""
        self._flush()
        if isinstance(records, np.ndarray):
            records = np.ascontiguousarray(records, dtype=np.uint8)
            ptr = records.ctypes.data
        else:
            ptr = _ptr(records)
        check(self._lib.r3d_tree_apply_delta_owned(self._h, ptr, int(n_records), int(part), int(nparts)), self._ctx.handle)

    BRICK_RECORD_BYTES = _lib.BRICK_RECORD_BYTES

    def numBricks(self):
        self._flush()
        n = C.c_uint64(0)
        check(self._lib.r3d_tree_num_bricks(self._h, C.byref(n)), self._ctx.handle)
        return int(n.value)

    def exportBricks(self, buf=None, capacity=None):
        """Every brick of the map as 2120-byte records.  With no buffer: returns a (n, 2120) uint8 host array."""
        self._flush()
        n = self.numBricks()
        cnt = C.c_uint64(0)
        if buf is None:
            rec = np.zeros((max(n, 1), _lib.BRICK_RECORD_BYTES), np.uint8)
            check(self._lib.r3d_tree_export_bricks(self._h, rec.ctypes.data, n, C.byref(cnt)), self._ctx.handle)
            return rec[:n]
        check(self._lib.r3d_tree_export_bricks(self._h, _ptr(buf), int(n if capacity is None else capacity), C.byref(cnt)), self._ctx.handle)
        return int(cnt.value)

    def importBricks(self, records, n_records=None):
        self._flush()
        if isinstance(records, np.ndarray):
            records = np.ascontiguousarray(records, dtype=np.uint8)
            n = records.size // _lib.BRICK_RECORD_BYTES if n_records is None else int(n_records)
            ptr = records.ctypes.data
        else:
            ptr, n = _ptr(records), int(n_records)
        check(self._lib.r3d_tree_import_bricks(self._h, ptr, n), self._ctx.handle)

    @staticmethod
    def deltaKeys(records):
        """Expand delta records to explicit OcTreeKeys: (free (n,3) uint16, occupied (m,3) uint16)."""
        lib = _lib.load()
        rec = np.ascontiguousarray(records, dtype=np.uint8)
        n = rec.size // _lib.DELTA_RECORD_BYTES
        nf, no = C.c_uint64(0), C.c_uint64(0)
        check(lib.r3d_delta_expand_keys(rec.ctypes.data, n, None, 0, C.byref(nf), None, 0, C.byref(no)))
        fk = np.zeros((max(nf.value, 1), 3), np.uint16)
        ok = np.zeros((max(no.value, 1), 3), np.uint16)
        check(lib.r3d_delta_expand_keys(rec.ctypes.data, n, fk.ctypes.data, nf.value, C.byref(nf), ok.ctypes.data, no.value, C.byref(no)))
        return fk[: nf.value], ok[: no.value]

    # ------------------------------------------------------------------ internals
    def _flush(self):
        if not self._pend_pts:
            return
        pts = np.array(self._pend_pts, dtype=np.float64).reshape(-1, 3).astype(np.float32)
        upd = np.array(self._pend_upd, dtype=np.float32)
        self._pend_pts, self._pend_upd = [], []
        # maximal runs of equal increments, applied in order
        cuts = np.flatnonzero(upd[1:] != upd[:-1]) + 1
        starts = np.concatenate([[0], cuts])
        ends = np.concatenate([cuts, [upd.size]])
        dropped = C.c_uint64(0)
        for a, b in zip(starts, ends):
            seg = np.ascontiguousarray(pts[a:b])
            check(self._lib.r3d_tree_update_points_logodds(self._h, seg.ctypes.data, seg.shape[0], float(upd[a]), C.byref(dropped)), self._ctx.handle)
            self.n_dropped += int(dropped.value)
