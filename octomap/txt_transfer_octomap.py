#!/usr/bin/env python3
"""Drop-in for the reference's octomap/txt_transfer_octomap.py: world `x,y,z` txt -> OcTree(0.1) -> .bt.

The reference is a module-level script with its paths edited in the source (:31-32); here the same two paths are
the positional arguments / FILE_TXT, FILE_BT constants.  `octomap` below is the GPU drop-in for the module the
reference imports (same OcTree / updateNode / updateInnerOccupancy / writeBinary calls).
"""
import argparse

from _bootstrap import package

octomap = package("octomap")
_m = package("mapping")

FILE_TXT = './point_world/5_22_31_changeyz_worldpoint.txt'
FILE_BT = './bt/airsim_5_22_31_changeyz_worldpoint.bt'
RESOLUTION = 0.1

txt_read = _m.txt_read


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("file_txt", nargs="?", default=FILE_TXT)
    ap.add_argument("file_bt", nargs="?", default=FILE_BT)
    ap.add_argument("--resolution", type=float, default=RESOLUTION)
    ap.add_argument("--device", type=int, default=0)
    a = ap.parse_args(argv)
    package("formats").ensure_dir(a.file_bt)
    tree = octomap.OcTree(a.resolution, device=a.device)
    txt_read(a.file_txt, tree)
    tree.updateInnerOccupancy()
    tree.writeBinary(bytes(a.file_bt, encoding='utf-8'))


if __name__ == '__main__':
    main()
