#!/usr/bin/env python3
"""North-star mode (BASELINE.json configs 3-5; no counterpart script in the reference, whose OctoMap scripts only call
updateNode): pose file + depth / disparity PNGs -> fused back-projection -> insertPointCloud per frame (ray-cast free
cells, endpoint occupied, clamped log-odds) -> .bt.  Same pose / depth conventions as transfer/camera_to_world.py."""
import argparse

from _bootstrap import package

_m = package("mapping")
_l = package("_lib")


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--qt-path", default='./camera_pose/image_colmap_simi_2.txt')
    ap.add_argument("--pose-format", choices=["comma", "colmap"], default="comma")
    ap.add_argument("--depth-dir", default='./depth/')
    ap.add_argument("--file-bt", default='./bt/sequence.bt')
    ap.add_argument("--resolution", type=float, default=0.1)
    ap.add_argument("--maxrange", type=float, default=80.0)
    ap.add_argument("--intrinsics", type=float, nargs=4, metavar=("FX", "FY", "CX", "CY"), default=[600.391, 600.079, 320, 240])
    ap.add_argument("--raw-depth", action="store_true", help="keep 16-bit samples (IMREAD_UNCHANGED) instead of IMREAD_GRAYSCALE")
    ap.add_argument("--depth-scale", type=float, default=1.0, help="metres (or disparity pixels) per raw unit")
    ap.add_argument("--disparity", action="store_true", help="samples are disparities: Z = fx*B/d")
    ap.add_argument("--baseline", type=float, default=0.25, help="stereo baseline B in metres (disparity mode)")
    ap.add_argument("--keep-invalid", action="store_true",
                    help="literal behaviour of the reference's back-projection: pixels with non-positive depth / disparity become points AT the "
                         "camera centre (the sensor's own voxel is then marked occupied in every frame); default: such pixels are dropped on the device")
    ap.add_argument("--drop-invalid", action="store_true", help="(default since round 2; kept for old command lines)")
    ap.add_argument("--device", type=int, default=0)
    a = ap.parse_args(argv)
    _m.DEVICE = a.device
    _m.pose_sequence_to_bt(a.qt_path, a.file_bt, a.intrinsics, depth_dir=a.depth_dir, resolution=a.resolution, maxrange=a.maxrange,
                           pose_format=a.pose_format, raw_depth=a.raw_depth, depth_scale=a.depth_scale,
                           mode=_l.MODE_DISPARITY if a.disparity else _l.MODE_DEPTH, fB=a.intrinsics[0] * a.baseline,
                           drop_invalid=not a.keep_invalid)


if __name__ == '__main__':
    main()
