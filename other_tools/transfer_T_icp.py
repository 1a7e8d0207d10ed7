#!/usr/bin/env python3
"""Drop-in for the reference's other_tools/transfer_T_icp.py: apply the 4x4 ICP transform in T_data.txt to
./point/24.txt, concatenate with ./point/0.txt, write ./point_world/03_testT.txt and ./ply/icp/024.ply
(defaults = the constants at transfer_T_icp.py:99-110).  The per-point product runs on the GPU."""
import argparse

from _bootstrap import package

_i = package("icp")

get_T = _i.get_T
point_camera = _i.point_camera
local_world = _i.local_world
genply = _i.genply


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--path-T", default='T_data.txt')
    ap.add_argument("--path-world", default='./point_world/03_testT.txt')
    ap.add_argument("--path-ply", default='./ply/icp/024.ply')
    ap.add_argument("--fixed", default='./point/0.txt')
    ap.add_argument("--moving", default='./point/24.txt')
    ap.add_argument("--device", type=int, default=0)
    a = ap.parse_args(argv)
    _i.DEVICE = a.device
    _i.run(a.path_T, a.path_world, a.path_ply, a.fixed, a.moving)


if __name__ == '__main__':
    main()
