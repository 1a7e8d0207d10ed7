"""a1: the native PNG decoder (host thread pool, no GPU needed) against OpenCV -- which is what the reference calls
(cv.imread(..., IMREAD_GRAYSCALE) at transfer/camera_to_world.py:160, IMREAD_UNCHANGED + [:, :, 1] at
transfer/pixel_to_camera.py:133-134).  Bit-exact for every PNG flavour a depth / disparity map comes in."""
import importlib
import os
import struct
import zlib

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
formats = importlib.import_module("3d_reconstruction_system_b200.formats")


def write_png_raw(path, W, H, depth, color, rows, filters=None, palette=None, chunk=None):
    """Minimal PNG writer for the flavours cv2.imwrite cannot produce (sub-byte grey, grey+alpha, palette, chosen filters,
    IDAT split in several chunks).  rows: H byte strings of un-filtered scanline data."""
    def ch(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xffffffff)
    bpp = max(1, {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}[color] * depth // 8)
    raw = bytearray()
    prev = bytes(len(rows[0]))
    for y, r in enumerate(rows):
        ft = (filters[y % len(filters)] if filters else 0)
        out = bytearray(len(r))
        for i in range(len(r)):
            a = r[i - bpp] if i >= bpp else 0
            b = prev[i]
            c = prev[i - bpp] if i >= bpp else 0
            if ft == 0: p = 0
            elif ft == 1: p = a
            elif ft == 2: p = b
            elif ft == 3: p = (a + b) >> 1
            else:
                pp = a + b - c
                pa, pb, pc = abs(pp - a), abs(pp - b), abs(pp - c)
                p = a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)
            out[i] = (r[i] - p) & 255
        raw.append(ft)
        raw += out
        prev = r
    z = zlib.compress(bytes(raw), 6)
    body = b"\x89PNG\r\n\x1a\n" + ch(b"IHDR", struct.pack(">IIBBBBB", W, H, depth, color, 0, 0, 0))
    if palette is not None:
        body += ch(b"PLTE", bytes(palette))
    if chunk:
        for i in range(0, len(z), chunk):
            body += ch(b"IDAT", z[i:i + chunk])
    else:
        body += ch(b"IDAT", z)
    body += ch(b"IEND", b"")
    with open(path, "wb") as f:
        f.write(body)


def check_all_modes(path):
    g = cv2.imread(path, cv2.IMREAD_GRAYSCALE)
    u = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    assert g is not None and u is not None
    assert np.array_equal(formats.imread_gray(path), g)
    w, h, c, d = formats.png_info(path)
    assert (h, w) == g.shape and c == (1 if u.ndim == 2 else u.shape[2]) and d == 8 * u.dtype.itemsize
    raw = formats.imread_raw(path)
    assert raw.dtype == u.dtype and np.array_equal(raw, u if u.ndim == 2 else u[:, :, 0])
    if u.ndim == 3:
        for k in range(u.shape[2]):
            got = formats.imread_batch([path], "channel", channel=k)[0]
            assert got.dtype == u.dtype and np.array_equal(got, u[:, :, k]), k
        assert np.array_equal(formats.imread_unchanged_green(path), u[:, :, 1])
    else:
        with pytest.raises(IndexError):
            formats.imread_unchanged_green(path)          # the reference's gt[:, :, 1] raises IndexError on a 2-D image


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16])
@pytest.mark.parametrize("channels", [1, 3, 4])
def test_cv2_written_files(tmp_path, dtype, channels):
    rng = np.random.default_rng(channels * 10 + np.dtype(dtype).itemsize)
    hi = 256 if dtype == np.uint8 else 65536
    for (H, W) in ((375, 1242), (7, 11), (1, 1), (33, 1025)):
        shape = (H, W) if channels == 1 else (H, W, channels)
        img = rng.integers(0, hi, size=shape).astype(dtype)
        if H > 100:      # smooth content exercises every adaptive filter type
            yy, xx = np.mgrid[0:H, 0:W]
            base = (xx * 3 + yy * 5) % hi
            img = ((base if channels == 1 else np.stack([base] * channels, axis=2)) + rng.integers(0, 4, size=shape)).astype(dtype)
        p = str(tmp_path / ("a_%d_%d.png" % (H, W)))
        assert cv2.imwrite(p, img)
        check_all_modes(p)


def test_hand_written_flavours(tmp_path):
    rng = np.random.default_rng(3)
    W, H = 37, 19
    # grey 1 / 2 / 4 bits, every filter type, IDAT split into small chunks
    for depth in (1, 2, 4):
        rb = (W * depth + 7) // 8
        rows = [bytes(rng.integers(0, 256, size=rb, dtype=np.uint8)) for _ in range(H)]
        p = str(tmp_path / ("g%d.png" % depth))
        write_png_raw(p, W, H, depth, 0, rows, filters=[0, 1, 2, 3, 4], chunk=23)
        check_all_modes(p)
    # grey + alpha, 8 and 16 bit
    for depth in (8, 16):
        rows = [bytes(rng.integers(0, 256, size=W * 2 * depth // 8, dtype=np.uint8)) for _ in range(H)]
        p = str(tmp_path / ("ga%d.png" % depth))
        write_png_raw(p, W, H, depth, 4, rows, filters=[4, 3, 1, 2, 0])
        check_all_modes(p)
    # palette, 8 and 4 bit indices
    pal = rng.integers(0, 256, size=256 * 3, dtype=np.uint8)
    rows = [bytes(rng.integers(0, 256, size=W, dtype=np.uint8)) for _ in range(H)]
    p = str(tmp_path / "pal8.png")
    write_png_raw(p, W, H, 8, 3, rows, filters=[1, 4], palette=pal)
    check_all_modes(p)
    rows = [bytes(rng.integers(0, 256, size=(W * 4 + 7) // 8, dtype=np.uint8)) for _ in range(H)]
    p = str(tmp_path / "pal4.png")
    write_png_raw(p, W, H, 4, 3, rows, palette=pal[:48])
    check_all_modes(p)
    # RGB / RGBA 16 bit with Paeth and Average rows
    for color, nc in ((2, 3), (6, 4)):
        rows = [bytes(rng.integers(0, 256, size=W * nc * 2, dtype=np.uint8)) for _ in range(H)]
        p = str(tmp_path / ("c%d.png" % color))
        write_png_raw(p, W, H, 16, color, rows, filters=[3, 4])
        check_all_modes(p)


def test_batch_threads_and_errors(tmp_path):
    rng = np.random.default_rng(9)
    frames = rng.integers(0, 65536, size=(40, 48, 64)).astype(np.uint16)
    paths = []
    for k in range(40):
        p = str(tmp_path / ("f%02d.png" % k))
        cv2.imwrite(p, frames[k])
        paths.append(p)
    for nt in (0, 1, 3, 64):
        assert np.array_equal(formats.imread_batch(paths, "raw", n_threads=nt), frames)
        assert np.array_equal(formats.imread_batch(paths, "gray", n_threads=nt), (frames >> 8).astype(np.uint8))
    out = np.empty((40, 48, 64), np.uint16)
    assert formats.imread_batch(paths, "raw", out=out) is out and np.array_equal(out, frames)
    with pytest.raises(ValueError):
        formats.imread_batch(paths, "raw", out=np.empty((40, 48, 64), np.uint8))
    # wrong size in the middle of a batch, a missing file, garbage, an interlaced file
    cv2.imwrite(str(tmp_path / "small.png"), frames[0][:10, :10])
    with pytest.raises(ValueError, match="expected 64x48"):
        formats.imread_batch(paths[:3] + [str(tmp_path / "small.png")] + paths[3:], "raw")
    with pytest.raises(FileNotFoundError):
        formats.imread_batch(paths[:2] + [str(tmp_path / "nope.png")], "raw")
    (tmp_path / "junk.png").write_bytes(b"\x89PNG\r\n\x1a\n" + b"\0" * 64)
    with pytest.raises(ValueError):
        formats.imread_batch([str(tmp_path / "junk.png")], "gray")
    assert formats.imread_batch([], "gray").shape[0] == 0


def _png_chunks(ihdr, idat):
    def ch(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xffffffff)
    return b"\x89PNG\r\n\x1a\n" + ch(b"IHDR", ihdr) + ch(b"IDAT", idat) + ch(b"IEND", b"")


def test_png_variants_only_opencv_reads_fall_back(tmp_path):
    """every file cv.imread (the reference's reader) accepts must be accepted.  Adam7-interlaced PNGs and streams
    with bytes after the image data (libpng only warns) are not served by the native decoder: the readers hand exactly those
    files to OpenCV, per file in imread_*, and by ending the native prefix in read_frame_batch."""
    rng = np.random.default_rng(9)
    H, W = 13, 21
    img = rng.integers(0, 65536, size=(H, W)).astype(np.uint16)
    # Adam7: seven reduced images, each as un-filtered scanlines
    raw = bytearray()
    for x0, y0, dx, dy in ((0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2)):
        sub = img[y0::dy, x0::dx]
        if sub.size == 0:
            continue
        for row in sub:
            raw.append(0)
            raw += row.astype(">u2").tobytes()
    inter = tmp_path / "interlaced.png"
    inter.write_bytes(_png_chunks(struct.pack(">IIBBBBB", W, H, 16, 0, 0, 0, 1), zlib.compress(bytes(raw))))
    ref = cv2.imread(str(inter), cv2.IMREAD_UNCHANGED)
    assert ref is not None and np.array_equal(ref, img)                       # OpenCV de-interlaces
    assert np.array_equal(formats.imread_raw(str(inter)), img)
    assert np.array_equal(formats.imread_gray(str(inter)), cv2.imread(str(inter), cv2.IMREAD_GRAYSCALE))
    # trailing bytes after the zlib stream inside IDAT
    plain = bytearray()
    for row in img:
        plain.append(0)
        plain += row.astype(">u2").tobytes()
    trail = tmp_path / "trailing.png"
    trail.write_bytes(_png_chunks(struct.pack(">IIBBBBB", W, H, 16, 0, 0, 0, 0), zlib.compress(bytes(plain)) + b"\x00" * 7))
    ref = cv2.imread(str(trail), cv2.IMREAD_UNCHANGED)
    if ref is not None:                                                       # (whatever OpenCV says goes)
        assert np.array_equal(formats.imread_raw(str(trail)), ref)
    # a batch: native prefix, then the interlaced file on its own, then native again
    good = []
    for k in range(3):
        p = tmp_path / ("g%d.png" % k)
        cv2.imwrite(str(p), np.roll(img, k, axis=1))
        good.append(str(p))
    paths = good[:2] + [str(inter)] + good[2:]
    stack, used = formats.read_frame_batch(paths, "raw")
    assert used == 2 and np.array_equal(stack[1], np.roll(img, 1, axis=1))
    stack, used = formats.read_frame_batch(paths[2:], "raw")
    assert used == 1 and np.array_equal(stack[0], img)
    stack, used = formats.read_frame_batch(paths[3:], "raw")
    assert used == 1 and np.array_equal(stack[0], np.roll(img, 2, axis=1))
    with pytest.raises((ValueError, FileNotFoundError)):
        bad = tmp_path / "broken.png"
        bad.write_bytes(_png_chunks(struct.pack(">IIBBBBB", W, H, 16, 0, 0, 0, 0), b"not a zlib stream"))
        formats.imread_raw(str(bad))
