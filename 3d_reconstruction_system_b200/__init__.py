"""B200-native mapping hot path of rainfall1998/3D_reconstruction_system.

depth / disparity frames -> camera-frame points -> world-frame points (PLY / txt) -> OctoMap occupancy (.bt),
computed by hand-written CUDA kernels for sm_100a behind the C ABI of include/r3d.h.  The Python modules
here mirror the reference's function / module interfaces:

    transfer      gentxtcord, scipy_transfer, point_camera, get_pointdata, genply*, get_file_name
    octomap       OcTree(res).updateNode / insertPointCloud / updateInnerOccupancy / writeBinary
    runtime       Context (one per GPU): batched back-projection, pose tables
    formats       pose files (comma format + Colmap images.txt), PLY / txt readers and writers

The package name is not a Python identifier; import it with
    r3d = importlib.import_module("3d_reconstruction_system_b200")
"""
__version__ = "0.1.0"

from . import _lib  # noqa: F401
from ._lib import MODE_DEPTH, MODE_DISPARITY, R3DError  # noqa: F401
from .runtime import Context, default_context  # noqa: F401
