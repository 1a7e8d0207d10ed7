"""CPU-side checks of the boundary: the shared library loads, exports every symbol include/r3d.h declares,
fails loudly without a GPU, and its host-side helpers (pose table, file formats) match the reference."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

from oracle import points_oracle as po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "r3d.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(r3d_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(r3d):
    lib = r3d._lib.load()
    names = _declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), "missing export " + n
    assert set(names) == set(r3d._lib.SIGNATURES), "ctypes table and header disagree"
    assert b"sm_100a" in lib.r3d_version()


def test_no_cpu_fallback_without_gpu(r3d):
    lib = r3d._lib.load()
    if lib.r3d_device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(r3d.R3DError, match="no CPU fallback"):
        r3d.Context(0)


def test_pose_to_rt_matches_scipy_transfer(r3d, golden_dir):
    lib = r3d._lib.load()
    rng = np.random.default_rng(7)
    q = rng.normal(size=(200, 4)) * rng.uniform(0.1, 5.0, size=(200, 1))   # non-unit on purpose
    t = rng.uniform(-2000, 2000, size=(200, 3))
    poses = np.ascontiguousarray(np.concatenate([q, t], axis=1))
    rt = np.zeros((200, 12))
    assert lib.r3d_pose_to_rt(poses.ctypes.data, 200, 1.0, rt.ctypes.data) == 0
    for k in range(200):
        ref = po.scipy_transfer(q[k])
        assert np.max(np.abs(rt[k, :9].reshape(3, 3) - ref)) <= 8 * np.finfo(np.float64).eps
        assert np.array_equal(rt[k, :9].reshape(3, 3), po.quat_to_rinv_fixed(q[k]))
        assert np.array_equal(rt[k, 9:], t[k])
    z = np.load(os.path.join(golden_dir, "ref_c2w_small.npz"))
    poses = np.ascontiguousarray(np.concatenate([z["quats"], z["trans"]], axis=1))
    rt = np.zeros((3, 12))
    assert lib.r3d_pose_to_rt(poses.ctypes.data, 3, 1.0, rt.ctypes.data) == 0
    assert np.max(np.abs(rt[:, :9].reshape(3, 3, 3) - z["rinv"])) <= 4 * np.finfo(np.float64).eps
    # ICP scale folded into t
    assert lib.r3d_pose_to_rt(poses.ctypes.data, 3, 2.5, rt.ctypes.data) == 0
    assert np.array_equal(rt[:, 9:], 2.5 * z["trans"])
    # scipy raises on a zero quaternion; so do we
    bad = np.zeros((1, 7))
    assert lib.r3d_pose_to_rt(bad.ctypes.data, 1, 1.0, rt.ctypes.data) == -1
    assert b"zero norm" in lib.r3d_last_error(None)


def test_ply_ascii_bytes_match_reference(r3d, golden_dir, tmp_path):
    from importlib import import_module
    formats = import_module("3d_reconstruction_system_b200.formats")
    z = np.load(os.path.join(golden_dir, "ref_c2w_small.npz"))
    txt = json.load(open(os.path.join(golden_dir, "ref_c2w_small_text.json")))
    w = z["world"].reshape(-1, 3)
    assert formats.ply_ascii_text(w[:, 0], w[:, 1], w[:, 2]) == txt["ply_txt"]
    p = tmp_path / "a.ply"
    formats.write_ply_ascii(str(p), w[:, 0], w[:, 1], w[:, 2])
    assert p.read_text() == txt["ply_txt"]
    # the reference's own PLY read back by ply_transfer_octomap.txt_read: 8 skipped lines drop the first vertex
    pts = formats.read_ply_points(str(p))
    assert pts.shape[0] == w.shape[0] - 1
    assert np.allclose(pts, np.round(w[1:], 4), atol=5.1e-5)
    assert formats.ply_ascii_text([], [], []) == po.genply_text([], [], [])
    rgb = np.arange(9).reshape(3, 3)
    t = formats.ply_ascii_text([1.0, 2, 3], [0.5, 0.25, 0.125], [-1, -2, -3.00005], rgb=rgb)
    assert "property uchar alpha" in t and "1.0000 0.5000 -1.0000 0 1 2 0\n" in t


def test_pose_file_and_txt_roundtrip(r3d, golden_dir, tmp_path):
    from importlib import import_module
    formats = import_module("3d_reconstruction_system_b200.formats")
    z = np.load(os.path.join(golden_dir, "ref_c2w_small.npz"))
    txt = json.load(open(os.path.join(golden_dir, "ref_c2w_small_text.json")))
    p = tmp_path / "pose.txt"
    p.write_text(txt["pose_txt"])
    poses = formats.read_pose_file(str(p))
    assert np.array_equal(poses["t"], z["trans"]) and np.array_equal(poses["q"], z["quats"])
    assert poses["names"] == ["f0.png", "f1.png", "f2.png"]
    # Colmap images.txt: scalar-first quaternion, every second line is 2-D points
    c = tmp_path / "images.txt"
    c.write_text("# Image list\n# comment\n1 0.9 0.1 0.2 0.3 0.5 -0.5 1.0 1 f0.png\n1.0 2.0 -1\n2 1 0 0 0 0 0 0 1 f1.png\n\n")
    cm = formats.read_colmap_images_txt(str(c))
    assert np.array_equal(cm["q"], np.array([[0.1, 0.2, 0.3, 0.9], [0, 0, 0, 1.0]]))
    assert cm["names"] == ["f0.png", "f1.png"]
    # world txt written with str() round-trips exactly
    w = z["world"][-1]
    q = tmp_path / "w.txt"
    formats.write_xyz_txt(str(q), w[:, 0], w[:, 1], w[:, 2])
    assert q.read_text() == txt["world_txt_last"]
    assert np.array_equal(formats.read_xyz_txt(str(q)), w)
    # camera txt prints the raw integer depth
    X, Y, Z = po.backproject(po.raw_to_z(z["depths"][0]))
    formats.write_xyz_txt(str(q), X, Y, Z, z_raw=z["depths"][0])
    assert q.read_text() == txt["cam_txt"][0]
