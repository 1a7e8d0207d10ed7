"""a8: the native point-cloud text readers (host thread pool, no GPU) against Python's own parsing -- what the reference's
txt_read loops do with float() (octomap/txt_transfer_octomap.py:16-28, octomap/ply_transfer_octomap.py:16-40)."""
import importlib
import time

import numpy as np
import pytest

from _cases import repr_cases

formats = importlib.import_module("3d_reconstruction_system_b200.formats")


def ref_txt(text):
    return np.array([[float(v) for v in ln.split(",")[:3]] for ln in text.splitlines() if ln.strip()], dtype=np.float64).reshape(-1, 3)


def ref_ply(text, skip=8, cap=5400001):
    rows = [ln.split() for ln in text.split("\n")[skip:]]
    return np.array([[float(v) for v in r[:3]] for r in rows if len(r) >= 3][:cap], dtype=np.float64).reshape(-1, 3)


def test_txt_exact_round_trip_of_hard_doubles(tmp_path):
    v = repr_cases(n=60000)
    v = v[np.isfinite(v)]
    x = v[: (v.size // 3) * 3].reshape(-1, 3)
    text = "".join("%r,%r,%r\n" % (a, b, c) for a, b, c in x.tolist())
    p = tmp_path / "w.txt"
    p.write_text(text)
    got = formats.read_xyz_txt(str(p))
    assert got.shape == x.shape and np.array_equal(got.view(np.uint64), x.view(np.uint64))      # bit-exact, file order kept


def test_txt_flavours(tmp_path):
    text = ("1,2,3\n"                       # integers (gentxtcord prints the raw depth as an integer)
            "-0.0,-0.0,0\n"
            " 1.5 , 2.5e-3 ,\t-7 \n"        # blanks around fields, like float(' 1.5 ')
            "1e400,-1e400,nan\n"            # overflow -> inf like float()
            "4,5,6,7,8\n"                   # extra columns ignored
            "\n"
            "9,10,11")                      # no trailing newline
    p = tmp_path / "a.txt"
    p.write_text(text)
    got = formats.read_xyz_txt(str(p))
    want = ref_txt(text)
    assert got.shape == want.shape == (6, 3)
    assert np.array_equal(np.nan_to_num(got, nan=123.0), np.nan_to_num(want, nan=123.0))
    assert np.signbit(got[1, 0]) and np.isinf(got[3, 0]) and np.isnan(got[3, 2])
    (tmp_path / "empty.txt").write_text("")
    assert formats.read_xyz_txt(str(tmp_path / "empty.txt")).shape == (0, 3)
    with pytest.raises(FileNotFoundError):
        formats.read_xyz_txt(str(tmp_path / "missing.txt"))


def test_ply_reader_semantics(tmp_path):
    rng = np.random.default_rng(1)
    pts = rng.normal(scale=300.0, size=(20000, 3))
    own = formats.ply_ascii_text(pts[:, 0], pts[:, 1], pts[:, 2])            # the reference writer's 7-line header + indentation
    p = tmp_path / "own.ply"
    p.write_text(own)
    got = formats.read_ply_points(str(p))
    want = ref_ply(own)
    assert got.shape == want.shape == (19999, 3) and np.array_equal(got, want)   # "skip 8" drops the first vertex, trailer skipped
    assert np.array_equal(formats.read_ply_points(str(p), skip_lines=7), ref_ply(own, 7))
    assert np.array_equal(formats.read_ply_points(str(p), max_points=1001), want[:1001])
    rgb = "ply\nformat ascii 1.0\ncomment x\nelement vertex 3\nproperty float x\nproperty float y\nproperty float z\nend_header\n1 2 3 255 0 0 0\n4.5 5.5 6.5 1 2 3 0\n\n7 8 9\n"
    q = tmp_path / "rgb.ply"
    q.write_text(rgb)
    assert np.array_equal(formats.read_ply_points(str(q)), np.array([[1, 2, 3], [4.5, 5.5, 6.5], [7, 8, 9.0]]))


def test_large_file_threads_keep_order(tmp_path):
    rng = np.random.default_rng(2)
    pts = np.round(rng.normal(scale=50.0, size=(400000, 3)), 4)
    p = tmp_path / "big.txt"
    with open(p, "w") as f:
        f.write("".join("%r,%r,%r\n" % t for t in map(tuple, pts.tolist())))
    t0 = time.perf_counter()
    got = formats.read_xyz_txt(str(p))
    dt = time.perf_counter() - t0
    assert np.array_equal(got, pts)
    for nt in (1, 3):
        assert np.array_equal(formats._read_text_points(str(p), 0, True, 0, n_threads=nt), pts)
    assert np.array_equal(formats._read_text_points(str(p), 100, True, 250001), pts[100:250101])
    print("parsed %d lines in %.3f s" % (pts.shape[0], dt))
