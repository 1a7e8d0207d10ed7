"""ctypes binding of libr3d_b200.so (the C ABI declared in include/r3d.h).

There is no CPU fallback: if the library is missing or no CUDA device is present the product path
raises.  Build it with `python 3d_reconstruction_system_b200/build.py` (or __graft_entry__.build()).
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# R3D_LIB_PATH: load an experiment build (build.py, R3D_BUILD_TAG) instead of the product library
LIB_PATH = os.environ.get("R3D_LIB_PATH") or os.path.join(HERE, "libr3d_b200.so")

OK = 0
U8, U16, F32 = 0, 1, 2
MODE_DEPTH, MODE_DISPARITY = 0, 1
PNG_GRAY8, PNG_CHANNEL, PNG_RAW = 0, 1, 2
OUT_F32, OUT_F64 = 0, 1
DELTA_RECORD_BYTES = 136
BRICK_RECORD_BYTES = 2120

_vp, _sz, _u64, _dbl, _flt, _i32 = C.c_void_p, C.c_size_t, C.c_uint64, C.c_double, C.c_float, C.c_int
_u64p = C.POINTER(C.c_uint64)

# name -> (restype, argtypes): every symbol include/r3d.h declares
SIGNATURES = {
    "r3d_version": (C.c_char_p, []),
    "r3d_device_count": (_i32, []),
    "r3d_create": (_vp, [_i32]),
    "r3d_destroy": (None, [_vp]),
    "r3d_last_error": (C.c_char_p, [_vp]),
    "r3d_set_blocking": (_i32, [_vp, _i32]),
    "r3d_synchronize": (_i32, [_vp]),
    "r3d_stream": (_vp, [_vp]),
    "r3d_launch_count": (_u64, [_vp]),
    "r3d_last_kernel_ms": (_flt, [_vp]),
    "r3d_host_alloc": (_vp, [_sz]),
    "r3d_host_free": (None, [_vp]),
    "r3d_device_alloc": (_vp, [_vp, _sz]),
    "r3d_device_free": (None, [_vp, _vp]),
    "r3d_memcpy": (_i32, [_vp, _vp, _vp, _sz]),
    "r3d_png_info": (_i32, [C.c_char_p, C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32)]),
    "r3d_png_decode_batch": (_i32, [_vp, _i32, _i32, _i32, _vp, _sz, _i32, _i32, _i32, _i32, _vp]),
    "r3d_read_xyz_text": (_i32, [C.c_char_p, _i32, _i32, _u64, _vp, _u64, _u64p, _i32]),
    "r3d_pose_to_rt": (_i32, [_vp, _i32, _dbl, _vp]),
    "r3d_backproject_rt": (_i32, [_vp, _vp, _i32, _i32, _i32, _sz, _i32, _vp, _vp, _i32, _dbl, _dbl, _i32, _i32, _vp, _vp]),
    "r3d_backproject": (_i32, [_vp, _vp, _i32, _i32, _i32, _sz, _i32, _vp, _vp, _i32, _dbl, _dbl, _i32, _vp, _vp]),
    "r3d_transform_points": (_i32, [_vp, _vp, _u64, _vp, _vp]),
    "r3d_pose_apply_points": (_i32, [_vp, _vp, _u64, _vp, _vp]),
    "r3d_inflate": (_i32, [_vp, _sz, _vp, _sz]),
    "r3d_format_ply_rows": (_i32, [_vp, _vp, _vp, _vp, _sz, _u64, _vp, _vp, _sz, C.POINTER(_sz)]),
    "r3d_format_txt_rows": (_i32, [_vp, _vp, _vp, _vp, _sz, _u64, _i32, _vp, _sz, C.POINTER(_sz)]),
    "r3d_tree_create": (_i32, [_vp, _dbl, C.POINTER(_vp)]),
    "r3d_tree_destroy": (None, [_vp]),
    "r3d_tree_clear": (_i32, [_vp]),
    "r3d_tree_reserve": (_i32, [_vp, _u64]),
    "r3d_tree_params": (_i32, [_vp, _vp]),
    "r3d_tree_resolution": (_i32, [_vp, C.POINTER(_dbl)]),
    "r3d_tree_update_points": (_i32, [_vp, _vp, _u64, _i32, _u64p]),
    "r3d_tree_update_points_f64": (_i32, [_vp, _vp, _u64, _i32, _u64p]),
    "r3d_tree_update_points_logodds": (_i32, [_vp, _vp, _u64, _flt, _u64p]),
    "r3d_tree_insert_scan": (_i32, [_vp, _vp, _u64, _vp, _dbl, _i32]),
    "r3d_tree_insert_scans": (_i32, [_vp, _vp, _vp, _vp, C.c_uint32, _dbl, _i32]),
    "r3d_scan_deltas_compute": (_i32, [_vp, _vp, _vp, _vp, C.c_uint32, _dbl, _i32, _vp, _u64, _vp]),
    "r3d_tree_apply_deltas_owned": (_i32, [_vp, _vp, _vp, C.c_uint32, C.c_uint32, C.c_uint32]),
    "r3d_tree_defer_deltas_owned": (_i32, [_vp, _vp, _vp, C.c_uint32, C.c_uint32, C.c_uint32]),
    "r3d_tree_flush_deferred": (_i32, [_vp]),
    "r3d_scan_delta_compute": (_i32, [_vp, _vp, _u64, _vp, _dbl, _i32, _u64p]),
    "r3d_scan_delta_export": (_i32, [_vp, _vp, _u64, _u64p]),
    "r3d_tree_apply_delta": (_i32, [_vp, _vp, _u64]),
    "r3d_tree_apply_delta_owned": (_i32, [_vp, _vp, _u64, C.c_uint32, C.c_uint32]),
    "r3d_tree_num_bricks": (_i32, [_vp, _u64p]),
    "r3d_tree_export_bricks": (_i32, [_vp, _vp, _u64, _u64p]),
    "r3d_tree_import_bricks": (_i32, [_vp, _vp, _u64]),
    "r3d_delta_expand_keys": (_i32, [_vp, _u64, _vp, _u64, _u64p, _vp, _u64, _u64p]),
    "r3d_tree_last_scan_stats": (_i32, [_vp, _vp]),
    "r3d_tree_pipeline_stats": (_i32, [_vp, _vp]),
    "r3d_tree_growth_stats": (_i32, [_vp, _vp]),
    "r3d_tree_update_inner_occupancy": (_i32, [_vp]),
    "r3d_tree_write_bt": (_i32, [_vp, C.c_char_p]),
    "r3d_tree_write_bt_mem": (_i32, [_vp, _vp, _sz, C.POINTER(_sz)]),
    "r3d_tree_to_max_likelihood": (_i32, [_vp]),
    "r3d_tree_read_bt": (_i32, [_vp, C.c_char_p]),
    "r3d_tree_read_bt_mem": (_i32, [_vp, _vp, _sz]),
    "r3d_tree_write_ot": (_i32, [_vp, C.c_char_p]),
    "r3d_tree_write_ot_mem": (_i32, [_vp, _vp, _sz, C.POINTER(_sz)]),
    "r3d_tree_read_ot": (_i32, [_vp, C.c_char_p]),
    "r3d_tree_read_ot_mem": (_i32, [_vp, _vp, _sz]),
    "r3d_tree_num_voxels": (_i32, [_vp, _u64p]),
    "r3d_tree_size": (_i32, [_vp, _u64p]),
    "r3d_tree_search": (_i32, [_vp, _vp, _u64, _vp, _vp]),
    "r3d_tree_export_voxels": (_i32, [_vp, _vp, _vp, _u64, _u64p]),
    "r3d_coord_to_key": (_i32, [_vp, _vp, _u64, _vp, _vp]),
}

_lib = None


def load():
    """Load libr3d_b200.so and type every entry point.  Raises if the extension is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libr3d_b200.so is not built (%s). Run `python 3d_reconstruction_system_b200/build.py`; "
            "this package has no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class R3DError(RuntimeError):
    pass


def check(rc, ctx=None):
    if rc != OK:
        msg = load().r3d_last_error(ctx)
        raise R3DError("r3d error %d: %s" % (rc, msg.decode("utf-8", "replace") if msg else "?"))
