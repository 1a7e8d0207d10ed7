"""The PNG decoder's own inflate (csrc/r3d_inflate.cu) against zlib: identical bytes for every valid stream, an error
(never a crash or a wrong answer) for truncated and corrupted ones.  CPU only: the library loads without a GPU."""
import ctypes as C
import importlib
import zlib

import numpy as np
import pytest

r3d = importlib.import_module("3d_reconstruction_system_b200")


@pytest.fixture(scope="module")
def lib():
    return r3d._lib.load()


def inflate(lib, stream, n):
    out = np.empty(max(n, 1) + 64, dtype=np.uint8)          # guard bytes behind the output
    out[:] = 0xA5
    src = np.frombuffer(stream, dtype=np.uint8) if len(stream) else np.zeros(1, np.uint8)
    rc = lib.r3d_inflate(src.ctypes.data, len(stream), out.ctypes.data, n)
    assert np.all(out[n:n + 64] == 0xA5) or n == 0 and np.all(out[1:] == 0xA5), "wrote past the output buffer"
    return rc, bytes(out[:n])


def payloads():
    rng = np.random.default_rng(2026)
    yield b""
    yield b"a"
    yield b"abc" * 1000                                      # overlapping matches, distance 3
    yield bytes(100000)                                      # distance 1 runs, maximum-length matches
    yield bytes(rng.integers(0, 256, 200000, dtype=np.uint8))                 # incompressible: stored blocks at level 0, literals otherwise
    yield bytes(rng.integers(0, 4, 300000, dtype=np.uint8))                   # short codes
    yield bytes((rng.integers(0, 256, 70000, dtype=np.uint8) // 37 * 37).astype(np.uint8))
    # depth-image like: 16-bit big-endian samples after PNG's Sub filter
    z = (np.cumsum(rng.integers(-3, 4, 465750)) + 20000).astype(">u2")
    d = np.frombuffer(z.tobytes(), dtype=np.uint8).astype(np.int16)
    sub = d.copy()
    sub[2:] = (d[2:] - d[:-2]) & 255
    yield bytes(sub.astype(np.uint8))
    # long-range matches (distances up to 32 K) and a skewed alphabet that needs sub-tables (codes longer than 11 bits)
    base = bytes(rng.integers(0, 256, 40000, dtype=np.uint8))
    yield base + base[:30000] + base[5000:35000]
    probs = 0.5 ** np.arange(1, 257)
    probs /= probs.sum()
    yield bytes(rng.choice(256, size=400000, p=probs).astype(np.uint8))
    yield bytes(rng.choice(256, size=3000, p=probs).astype(np.uint8))


@pytest.mark.parametrize("level", [0, 1, 6, 9])
def test_matches_zlib_on_valid_streams(lib, level):
    for data in payloads():
        for wbits in (15, 9):
            co = zlib.compressobj(level, zlib.DEFLATED, wbits)
            stream = co.compress(data) + co.flush()
            rc, got = inflate(lib, stream, len(data))
            assert rc == 0, (level, wbits, len(data), r3d._lib.last_error() if hasattr(r3d._lib, "last_error") else rc)
            assert got == data


def test_fixed_huffman_and_multi_block_streams(lib):
    rng = np.random.default_rng(7)
    data = bytes(rng.integers(0, 8, 50000, dtype=np.uint8))
    co = zlib.compressobj(6, zlib.DEFLATED, 15, 8, zlib.Z_FIXED)
    stream = co.compress(data) + co.flush()
    rc, got = inflate(lib, stream, len(data))
    assert rc == 0 and got == data
    # many blocks of mixed kinds: full flushes insert empty stored blocks, levels change between blocks
    co = zlib.compressobj(9)
    parts = []
    for i in range(40):
        chunk = bytes(rng.integers(0, 256 if i % 3 == 0 else 5, 1 + 997 * i, dtype=np.uint8))
        parts.append(chunk)
    stream = b"".join(co.compress(p) + co.flush(zlib.Z_FULL_FLUSH if i % 2 else zlib.Z_SYNC_FLUSH) for i, p in enumerate(parts)) + co.flush()
    data = b"".join(parts)
    rc, got = inflate(lib, stream, len(data))
    assert rc == 0 and got == data


def test_wrong_sizes_and_truncation_are_errors(lib):
    data = bytes(np.random.default_rng(3).integers(0, 50, 20000, dtype=np.uint8))
    stream = zlib.compress(data, 6)
    assert inflate(lib, stream, len(data))[0] == 0
    assert inflate(lib, stream, len(data) - 1)[0] != 0       # output does not fit
    assert inflate(lib, stream, len(data) + 1)[0] != 0       # output shorter than promised
    for cut in (0, 1, 2, 5, len(stream) // 2, len(stream) - 5, len(stream) - 1):
        assert inflate(lib, stream[:cut], len(data))[0] != 0
    bad = bytearray(stream)
    bad[-1] ^= 1                                             # Adler-32
    assert inflate(lib, bytes(bad), len(data))[0] != 0
    assert inflate(lib, b"\x78\x9d" + stream[2:], len(data))[0] != 0     # FCHECK
    assert inflate(lib, b"\x79\x9c" + stream[2:], len(data))[0] != 0     # CM != 8


def test_corrupted_streams_never_disagree_with_zlib(lib):
    """Bit flips anywhere in the stream: either both decoders reject it, or both produce the same bytes (a flip in a
    stored block's payload is caught by the Adler-32 in both)."""
    rng = np.random.default_rng(11)
    datas = [bytes(rng.integers(0, 30, 6000, dtype=np.uint8)), bytes(rng.integers(0, 256, 3000, dtype=np.uint8)), b"xyz" * 3000]
    agree_ok = agree_bad = 0
    for data in datas:
        for level in (0, 1, 9):
            stream = zlib.compress(data, level)
            for _ in range(300):
                bad = bytearray(stream)
                for _ in range(int(rng.integers(1, 4))):
                    bad[int(rng.integers(0, len(bad)))] ^= 1 << int(rng.integers(0, 8))
                bad = bytes(bad)
                try:
                    ref = zlib.decompress(bad)
                    if len(ref) != len(data):
                        ref = None
                except zlib.error:
                    ref = None
                rc, got = inflate(lib, bad, len(data))
                if ref is None:
                    assert rc != 0
                    agree_bad += 1
                else:
                    assert rc == 0 and got == ref
                    agree_ok += 1
    assert agree_bad > 1000
