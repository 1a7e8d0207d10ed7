#!/usr/bin/env python3
# -*- coding:utf-8 -*-
"""Drop-in for the reference's transfer/camera_to_world.py: pose file + depth PNGs -> world points -> one merged PLY.

Defaults are the reference's hard-coded ones (camera_to_world.py:68-71, 87, 160-163, 174, 179): pose file
./camera_pose/image_colmap_simi_2.txt (comma format: id,tx,ty,tz,qx,qy,qz,qw,name,...), depth PNGs in ./depth/
(IMREAD_GRAYSCALE), per-frame camera txt in ./point/, world txt ./point_world/small_worldpoint_5_23_5.txt, merged
ASCII PLY ./ply/small_035_p8.ply.  All frames of one image shape go through ONE fused GPU launch
(decode -> back-project -> pose transform); `--pose-format colmap` reads a raw Colmap images.txt instead.
"""
import argparse

from _bootstrap import package

_t = package("transfer")

str_tofloat = _t.str_tofloat
scipy_transfer = _t.scipy_transfer
point_camera = _t.point_camera
gentxtcord = _t.gentxtcord
get_pointdata = _t.get_pointdata
genply = _t.genply
get_file_name = _t.get_file_name


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--qt-path", default='./camera_pose/image_colmap_simi_2.txt')
    ap.add_argument("--pose-format", choices=["comma", "colmap"], default="comma")
    ap.add_argument("--depth-dir", default=_t.DEPTH_DIR)
    ap.add_argument("--point-dir", default=_t.POINT_DIR)
    ap.add_argument("--point-world-path", default=_t.POINT_WORLD_PATH)
    ap.add_argument("--ply-path", default=_t.PLY_PATH)
    ap.add_argument("--intrinsics", type=float, nargs=4, metavar=("FX", "FY", "CX", "CY"))
    ap.add_argument("--no-intermediate", action="store_true", help="skip the per-frame txt side files")
    ap.add_argument("--device", type=int, default=0)
    a = ap.parse_args(argv)
    _t.DEVICE = a.device
    _t.DEPTH_DIR, _t.POINT_DIR, _t.POINT_WORLD_PATH, _t.PLY_PATH = a.depth_dir, a.point_dir, a.point_world_path, a.ply_path
    get_file_name(a.qt_path, intr=a.intrinsics, write_intermediate=not a.no_intermediate, pose_format=a.pose_format)


if __name__ == '__main__':
    main()
