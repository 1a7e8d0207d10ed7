#!/usr/bin/env python3
"""Drop-in for the reference's octomap/ply_transfer_octomap.py (= other_tools/ply_transfer_octomap.py):
ASCII PLY -> OcTree(0.1) -> .bt, skipping exactly 8 header lines and stopping after point 5 400 000 like the
reference's txt_read (:16-40).  Blank / indentation-only rows are skipped instead of raising (documented deviation).
"""
import argparse

from _bootstrap import package

octomap = package("octomap")
_m = package("mapping")

FILE_PLY = './point/26_31_R-T.ply'
FILE_BT = './bt/airsim_26_31_R-T.bt'
RESOLUTION = 0.1
HEADER_LINES = 8
POINT_CAP = 5400001


def txt_read(file_path, tree):
    return _m.ply_read(file_path, tree, skip_lines=HEADER_LINES, max_points=POINT_CAP)


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("file_ply", nargs="?", default=FILE_PLY)
    ap.add_argument("file_bt", nargs="?", default=FILE_BT)
    ap.add_argument("--resolution", type=float, default=RESOLUTION)
    ap.add_argument("--header-lines", type=int, default=HEADER_LINES)
    ap.add_argument("--point-cap", type=int, default=POINT_CAP)
    ap.add_argument("--device", type=int, default=0)
    a = ap.parse_args(argv)
    package("formats").ensure_dir(a.file_bt)
    tree = octomap.OcTree(a.resolution, device=a.device)
    _m.ply_read(a.file_ply, tree, skip_lines=a.header_lines, max_points=a.point_cap)
    tree.updateInnerOccupancy()
    tree.writeBinary(bytes(a.file_bt, encoding='utf-8'))


if __name__ == '__main__':
    main()
