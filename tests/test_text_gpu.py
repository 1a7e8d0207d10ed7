"""K6: GPU "%.4f" PLY rows against Python's own formatter (which IS the reference's float_formatter,
transfer/camera_to_world.py:117) -- byte-exact, including ties, huge magnitudes, signed zeros, inf / nan."""
import importlib
import os

import numpy as np
import pytest

from _cases import fixed4_cases, repr_cases
from oracle import points_oracle as po

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(r3d):
    c = r3d.Context(0)
    yield c
    c.close()


def py_rows(x, rgb=None):
    if rgb is None:
        return "".join("%.4f %.4f %.4f \n" % (a, b, c) for a, b, c in x.tolist()).encode()
    return "".join("%.4f %.4f %.4f %d %d %d 0\n" % (a, b, c, r, g, bb) for (a, b, c), (r, g, bb) in zip(x.tolist(), rgb.tolist())).encode()


def test_rows_adversarial_values(ctx):
    v = fixed4_cases()
    x = v[: (v.size // 3) * 3].reshape(-1, 3).copy()
    assert ctx.ply_rows(x) == py_rows(x)
    # three separate arrays (the genply calling convention) and the rgb variant
    rgb = np.random.default_rng(5).integers(0, 256, size=(x.shape[0], 3), dtype=np.uint8)
    rgb[:3] = [[0, 9, 10], [99, 100, 255], [1, 2, 3]]
    assert ctx.ply_rows(x[:, 0].copy(), x[:, 1].copy(), x[:, 2].copy(), rgb=rgb) == py_rows(x, rgb)
    # a tile made only of astronomic values takes the row-by-row path
    big = np.full((700, 3), 1.2345678e300)
    big[::7] *= -1
    assert ctx.ply_rows(big) == py_rows(big)


@pytest.mark.parametrize("n", [0, 1, 255, 256, 257, 100003])
def test_rows_sizes_and_tile_boundaries(ctx, n):
    rng = np.random.default_rng(n)
    x = rng.normal(scale=200.0, size=(n, 3))
    assert ctx.ply_rows(x) == py_rows(x)


def test_million_points_and_device_input(ctx, r3d):
    rng = np.random.default_rng(1)
    x = rng.normal(scale=1500.0, size=(1 << 20, 3))
    want = py_rows(x)
    assert ctx.ply_rows(x) == want
    d = ctx.to_device(x)
    import ctypes as C
    need = C.c_size_t(0)
    lib = ctx.lib
    p = d.data_ptr()
    assert lib.r3d_format_ply_rows(ctx.handle, p, p + 8, p + 16, 3, x.shape[0], None, None, 0, C.byref(need)) == 0
    assert need.value == len(want)
    out = ctx.device_empty((need.value + 15) // 16 * 16, np.uint8)
    assert lib.r3d_format_ply_rows(ctx.handle, p, p + 8, p + 16, 3, x.shape[0], None, out.data_ptr(), need.value, C.byref(need)) == 0
    assert out.numpy()[: need.value].tobytes() == want


def test_genply_files_equal_reference_text(ctx, r3d, golden_dir, tmp_path):
    import json
    transfer = importlib.import_module("3d_reconstruction_system_b200.transfer")
    z = np.load(os.path.join(golden_dir, "ref_c2w_small.npz"))
    txt = json.load(open(os.path.join(golden_dir, "ref_c2w_small_text.json")))
    w = z["world"].reshape(-1, 3)
    p = tmp_path / "a.ply"
    transfer.genply([w[:, 0].tolist(), w[:, 1].tolist(), w[:, 2].tolist()], str(p), w.shape[0])
    assert p.read_text() == txt["ply_txt"]                     # the reference's own output
    transfer.genply_RGB([[], [], []], str(p))
    assert p.read_text() == po.genply_text([], [], [])
    with pytest.raises(ValueError):
        transfer.genply([[1.0], [2.0], [3.0]], str(p), 2)


def test_txt_rows_match_python_repr(ctx):
    v = repr_cases()
    x = v[: (v.size // 3) * 3].reshape(-1, 3).copy()
    want = "".join("%r,%r,%r\n" % (a, b, c) for a, b, c in x.tolist()).encode()
    assert ctx.txt_rows(x) == want
    assert ctx.txt_rows(x[:, 0].copy(), x[:, 1].copy(), x[:, 2].copy()) == want
    zi = np.random.default_rng(2).integers(0, 65536, size=x.shape[0])
    want = "".join("%r,%r,%d\n" % (a, b, c) for (a, b, _), c in zip(x.tolist(), zi.tolist())).encode()
    assert ctx.txt_rows(x[:, 0].copy(), x[:, 1].copy(), zi.astype(np.float64), z_is_integer=True) == want
    assert ctx.txt_rows(np.zeros((0, 3))) == b""


def test_gentxtcord_text_equals_reference(ctx, r3d, golden_dir, tmp_path):
    """gentxtcord through the GPU formatter writes the bytes the reference wrote (tests/golden: 640x480 frame, sha256)."""
    import hashlib
    import json
    transfer = importlib.import_module("3d_reconstruction_system_b200.transfer")
    z = np.load(os.path.join(golden_dir, "ref_p2c_640.npz"))
    meta = json.load(open(os.path.join(golden_dir, "ref_meta.json")))["p2c_640"]
    p = tmp_path / "p.txt"
    out = transfer.gentxtcord(str(p), z["depth"])
    txt = p.read_text()
    assert txt.splitlines()[:5] == meta["txt_first_lines"]
    assert hashlib.sha256(txt.encode()).hexdigest() == meta["txt_sha256"]
    assert np.array_equal(np.asarray(out[0])[z["sel"]], z["X"]) and np.array_equal(np.asarray(out[1])[z["sel"]], z["Y"])
