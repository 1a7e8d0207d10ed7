#!/usr/bin/env python3
# -*- coding:utf-8 -*-
"""Drop-in for the reference's transfer/pixel_to_camera.py: one depth PNG -> camera-frame `x,y,z` txt + ASCII PLY.

Same relative paths and constants as the reference's main() (pixel_to_camera.py:126-136): frame number 24,
./depth/24.png (IMREAD_UNCHANGED, green channel), ./point/24.txt, ./img/24.png, ./ply/24.ply, intrinsics
600.391 / 600.079 / 320 / 240.  The per-pixel loop runs on the GPU (libr3d_b200.so); there is no CPU fallback.
Every constant can be overridden on the command line.  The reference's main() ends in a TypeError (it passes three
arguments to the two-argument genply_RGB, :136); here `--rgb` selects the coloured writer the call was aiming at
(genply_noRGB, which despite its name writes rgb) and the default writes xyz only (genply_RGB).
"""
import argparse

from _bootstrap import package

_t = package("transfer")
_f = package("formats")

# the reference's importable names
gentxtcord = _t.gentxtcord
genply_noRGB = _t.genply_noRGB
genply_RGB = _t.genply_RGB


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--num", default="24", help="frame stem used by the default paths")
    ap.add_argument("--depth-path")
    ap.add_argument("--point-path")
    ap.add_argument("--img-path")
    ap.add_argument("--ply-path")
    ap.add_argument("--rgb", action="store_true", help="write x y z r g b 0 rows from --img-path (genply_noRGB)")
    ap.add_argument("--intrinsics", type=float, nargs=4, metavar=("FX", "FY", "CX", "CY"))
    ap.add_argument("--device", type=int, default=0)
    a = ap.parse_args(argv)
    _t.DEVICE = a.device
    num = a.num
    depth_path = a.depth_path or './depth/' + str(num) + '.png'
    point_path = a.point_path or './point/' + str(num) + '.txt'
    imgpath = a.img_path or './img/' + str(num) + '.png'
    pc_file = a.ply_path or './ply/' + str(num) + '.ply'
    gray_img = _f.imread_unchanged_green(depth_path)
    _f.ensure_dir(point_path)
    _f.ensure_dir(pc_file)
    gt_cord = gentxtcord(point_path, gray_img, intr=a.intrinsics)
    if a.rgb:
        genply_noRGB(gt_cord, imgpath, pc_file)
    else:
        genply_RGB(gt_cord, pc_file)


if __name__ == '__main__':
    main()
