"""Host-side mirror of other_tools/transfer_T_icp.py: apply a 4x4 ICP transform T to one `x,y,z` txt cloud,
concatenate it with an untouched one, write the merged txt and ASCII PLY.  The per-point product runs on the GPU
(r3d_transform_points); this module reads / writes the reference's files.

Reference                               here
  get_T(path_txt)          :33-43        get_T(path_txt)
  point_camera(p1, T)      :10-12        point_camera(p1, T)   (T . [x y z 1]^T; returns all 4 rows like np.dot does)
  local_world(...)         :71-97        local_world(path_local, file_write, T, xcord, ycord, zcord, flag)
  script body              :99-110       run(path_T, path_world, path_ply, path_fixed, path_moving)
"""
import numpy as np

from . import formats
from .runtime import default_context
from .transfer import genply  # same text as transfer_T_icp.py:46-68

DEVICE = 0


def str_tofloat(data):
    return np.array([float(v) for v in data])


def get_T(path_txt):
    """First four whitespace-separated rows of the file -> 4x4 float64 (transfer_T_icp.py:33-43)."""
    T = np.zeros((4, 4))
    with open(path_txt, 'r') as f:
        for i in range(4):
            row = str_tofloat(f.readline().split())
            T[i, 0:4] = row[0:4]
    return T


def transform_cloud(xyz, T):
    """(n,3) float64 -> first three rows of T . [x y z 1]^T per point, on the GPU."""
    return default_context(DEVICE).transform_points(xyz, T)


def point_camera(p1, T):
    """np.dot(T, p1.T).T for one homogeneous point (4,) or an (n,4) block; the fourth row is evaluated on the host
    (it is [0 0 0 1] . p for a rigid / similarity T and never written to a file)."""
    p = np.asarray(p1, dtype=np.float64)
    single = p.ndim == 1
    P = p.reshape(-1, 4)
    Tm = np.asarray(T, dtype=np.float64)
    if np.all(P[:, 3] == 1.0):
        xyz = transform_cloud(P[:, 0:3], Tm)
    else:   # general homogeneous coordinate: fold w into the translation column per point
        xyz = np.stack([transform_cloud(P[i:i + 1, 0:3], Tm @ np.diag([1, 1, 1, P[i, 3]]))[0] for i in range(P.shape[0])])
    w = ((Tm[3, 0] * P[:, 0] + Tm[3, 1] * P[:, 1]) + Tm[3, 2] * P[:, 2]) + Tm[3, 3] * P[:, 3]
    out = np.concatenate([xyz, w[:, None]], axis=1)
    return out[0] if single else out


def local_world(path_local, file_write, T, xcord, ycord, zcord, flag):
    """transfer_T_icp.py:71-97: read `x,y,z` lines; flag -> transform by T, else keep; append to the coordinate lists
    and write `str(x),str(y),str(z)` lines to the open file."""
    print('start transfer')
    pts = formats.read_xyz_txt(path_local)
    if flag and pts.shape[0]:
        pts = transform_cloud(pts, T)
    xcord.extend(pts[:, 0].tolist())
    ycord.extend(pts[:, 1].tolist())
    zcord.extend(pts[:, 2].tolist())
    rows = default_context(DEVICE).txt_rows(pts)          # str(x),str(y),str(z) lines, formatted on the GPU
    file_write.write(rows.decode("ascii"))


def run(path_T='T_data.txt', path_world='./point_world/03_testT.txt', path_ply='./ply/icp/024.ply',
        path_fixed='./point/0.txt', path_moving='./point/24.txt'):
    """Script body (transfer_T_icp.py:99-110)."""
    T = get_T(path_T)
    xcord, ycord, zcord = [], [], []
    formats.ensure_dir(path_world)
    with open(path_world, 'w') as file_w:
        local_world(path_fixed, file_w, T, xcord, ycord, zcord, False)
        local_world(path_moving, file_w, T, xcord, ycord, zcord, True)
    formats.ensure_dir(path_ply)
    genply([xcord, ycord, zcord], path_ply, len(xcord))
    return xcord, ycord, zcord
