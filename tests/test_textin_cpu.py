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


def _check_tokens(tmp_path, tokens, name):
    """Every token float() accepts three times on a line: bit-identical values.  Every token float() rejects: the reader
    raises ValueError naming the line, as the reference's str_tofloat does (no row is ever dropped silently)."""
    want, good, bad = [], [], []
    for tok in tokens:
        try:
            want.append(float(tok))
            good.append(tok)
        except ValueError:
            bad.append(tok)
    p = tmp_path / name
    p.write_text("".join("%s,%s,%s\n" % (t, t, t) for t in good))
    got = formats.read_xyz_txt(str(p))
    want = np.array(want, dtype=np.float64)
    assert got.shape == (want.size, 3)
    for c in range(3):
        col = np.ascontiguousarray(got[:, c])
        nan = np.isnan(want)
        assert np.array_equal(np.isnan(col), nan)
        assert np.array_equal(col[~nan].view(np.uint64), want[~nan].view(np.uint64)), name
    for k, tok in enumerate(bad):
        q = tmp_path / ("bad_%d_%s" % (k, name))
        q.write_text("1,2,3\n4,5,6\n7,%s,9\n10,11,12\n" % tok)
        with pytest.raises(ValueError, match="line 3"):
            formats.read_xyz_txt(str(q))


def test_malformed_rows_raise_like_the_reference(tmp_path):
    """a corrupted or truncated cloud must not lose points silently.  The reference's loops raise ValueError on
    any field float() rejects (str_tofloat converts EVERY field of the line) and fail on rows with fewer than three
    fields; only empty / blank lines are skipped here (the documented deviation)."""
    def txt(name, text):
        p = tmp_path / name
        p.write_text(text)
        return str(p)
    for k, (text, line) in enumerate([("1,2,3\n1,2,x\n4,5,6\n", 2), ("1,2,3\n1,2\n", 2), ("foo\n1,2,3\n", 1), ("1,2,3\n4,5,6\n1,2,0x10\n", 3),
                                      ("1,2,3,\n", 1), ("1,,3\n", 1), ("1,2,3,abc\n", 1), ("1,2,3\n4,5,6", 0), ("1,2,3\n   \n\t\n4,5,6\n\n", 0)]):
        if line:
            with pytest.raises(ValueError, match="line %d" % line):
                formats.read_xyz_txt(txt("t%d.txt" % k, text))
        else:
            assert np.array_equal(formats.read_xyz_txt(txt("t%d.txt" % k, text)), np.array([[1, 2, 3], [4, 5, 6.0]]))
    hdr = "h\n" * 8
    assert np.array_equal(formats.read_ply_points(txt("a.ply", hdr + "1 2 3\n4 5 6 255 0 0 0\n\n    ")), np.array([[1, 2, 3], [4, 5, 6.0]]))
    for k, (body, line) in enumerate([("1 2 3\n4 5\n", 10), ("1 2 3\n4 5 six\n", 10), ("1 2 3 r g b\n", 9), ("1,2,3\n", 9)]):
        with pytest.raises(ValueError, match="line %d" % line):
            formats.read_ply_points(txt("b%d.ply" % k, hdr + body))
    # the reference stops reading at its point cap: a malformed line after the cap is never looked at
    capped = txt("c.ply", hdr + "1 2 3\n4 5 6\n7 8 9\nbroken\n")
    assert formats.read_ply_points(capped, max_points=3).shape == (3, 3)
    with pytest.raises(ValueError, match="line 12"):
        formats.read_ply_points(capped, max_points=4)
    # many threads: the error names the FIRST malformed line of the file
    big = "".join("%d,%d,%d\n" % (i, i + 1, i + 2) for i in range(200000))
    lines = big.splitlines(True)
    lines[150000] = "1,2,oops\n"
    lines[60000] = "1,2\n"
    with pytest.raises(ValueError, match="line 60001"):
        formats.read_xyz_txt(txt("big.txt", "".join(lines)))


def test_decimal_to_double_is_correctly_rounded(tmp_path):
    """The readers' own decimal -> double conversion (csrc/r3d_strtod.cuh: Clinger's exact case + Eisel-Lemire, strtod for
    the rest) against Python's float() on what is hard for it: 17-19 digit significands, every decade of exponent, exact
    ties, the subnormal and overflow borders, long digit strings, odd but legal spellings."""
    rng = np.random.default_rng(123)
    toks = []
    # random significands of 1..19 digits with exponents over the whole range
    for _ in range(120000):
        nd = int(rng.integers(1, 20))
        m = int(rng.integers(10 ** (nd - 1), 10 ** nd, dtype=np.uint64)) if nd < 20 else 0
        q = int(rng.integers(-345, 310))
        toks.append("%de%d" % (m, q))
    # decimal points in every position, signs, exponent spellings
    for _ in range(40000):
        m = str(int(rng.integers(0, 10 ** 17)))
        k = int(rng.integers(0, len(m) + 1))
        s = m[:k] + "." + m[k:]
        if s == ".":
            s = "0."
        if rng.random() < 0.3:
            s += ("e", "E")[int(rng.integers(2))] + ("", "+", "-")[int(rng.integers(3))] + str(int(rng.integers(0, 40)))
        toks.append(("", "-", "+")[int(rng.integers(3))] + s)
    # exact ties and their neighbours around 2^53 .. 2^63 and small powers of ten
    for k in range(53, 64):
        for d in (-2, -1, 0, 1, 2, 3):
            toks.append(str((1 << k) + d))
            toks.append(str((1 << k) + (1 << (k - 53)) + d))          # half way between two doubles
    toks += ["9007199254740993", "9007199254740992.5", "9007199254740993.0000000000000001", "1e23", "8.41e21", "9.5e21", "5e-324", "4.9e-324",
             "2.4703282292062327e-324", "2.4703282292062328e-324", "2.2250738585072011e-308", "2.2250738585072014e-308", "1.7976931348623157e308",
             "1.7976931348623158e308", "1.7976931348623159e308", "1e309", "-1e309", "1e-400", "0e999999", "0.0e-999999", "1e999999", "1e-999999",
             "123456789012345678901234567890", "0.000000000000000000000000000000000012345678901234567890123",
             "1" + "0" * 40, "0." + "0" * 40 + "1", "3.14159265358979323846264338327950288", "000001.5", ".5", "5.", "+.5e1", "-0.0", "0", "00", "1e5",
             "1E-5", "1.e2", "inf", "-inf", "nan", "infinity", "Infinity", "NaN",
             # what float() rejects (the reader raises, like the reference)
             ".", "e5", "1e", "1e+", "--1", "1.5x", "1..5", "+-1", "0x10", "0x1p3", "nan(1)", "1__0", "_1", "1_", "1_.5", "1e_5", "in", "infinit",
             # PEP 515 underscores, which float() accepts between digits
             "1_000", "1_0.2_5e1_0", "1" * 70, "-" + "9" * 400 + ".5"]
    # the classic hard cases of decimal -> double conversion
    toks += ["6.0221409e+23", "1.0000000000000002", "1.00000000000000011102230246251565404236316680908203125",
             "1.00000000000000011102230246251565404236316680908203124", "1.00000000000000011102230246251565404236316680908203126",
             "0.1", "0.2", "0.3", "1.1", "2.2", "123456.789e3", "7.2057594037927933e16", "4.4501477170144023e-308", "4.503599627370497e15",
             "9.007199254740993e15", "3.237883913302901289588352412501532174863037669423108059901297049552301970670676565786835742587799557860615776559838283435514391084153169252689190564396459577394618038928365305143463955100356696665629202017331344031730044369360205258345803431471660032699580731300954848363975548690010751530018881758184174569652173110473696022749934638425380623369774736560008997404060967498028389191878963968575439222206416981462690113342524002724385941651051293552601421155333430225237291523843322331326138431477823591142408800030775170625915670728657003151953664260769822494937951845801530895238439819708403389937873241463484205608000027270531106827387907791444918534771598750162812548862768493201518991668028251730299953143924168545708663913329985436296e-308"]
    toks = [t for t in toks if "," not in t and "\n" not in t]
    _check_tokens(tmp_path, toks, "hard.txt")


def test_decimal_to_double_on_repr_and_fixed_notation(tmp_path):
    """What the files of the path contain: str(float64) fields (the txt files) and "%.4f" fields (the PLY rows)."""
    rng = np.random.default_rng(77)
    v = np.concatenate([rng.normal(scale=1e3, size=150000), rng.uniform(-1, 1, 50000) * 10.0 ** rng.integers(-300, 300, 50000),
                        np.frombuffer(rng.bytes(8 * 100000), dtype=np.float64)])
    v = v[np.isfinite(v)]
    _check_tokens(tmp_path, [repr(float(x)) for x in v], "repr.txt")
    _check_tokens(tmp_path, ["%.4f" % x for x in v[:150000]], "fixed.txt")
    _check_tokens(tmp_path, ["%.17e" % x for x in v[150000:250000]], "sci.txt")
