#!/usr/bin/env python3
"""Drop-in for the reference's other_tools/ply_transfer_octomap.py, which is byte-identical to its
octomap/ply_transfer_octomap.py: same entry point, same defaults, one implementation."""
import os
import runpy
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_TWIN = os.path.join(os.path.dirname(_HERE), "octomap", "ply_transfer_octomap.py")

if __name__ == '__main__':
    sys.path.insert(0, os.path.dirname(_TWIN))
    runpy.run_path(_TWIN, run_name="__main__")
