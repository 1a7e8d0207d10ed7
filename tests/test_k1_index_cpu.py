"""The pixel bookkeeping of the bulk K1 kernels (csrc/r3d_backproject.cu: k1_bulk_vec, k1_bulk_compact, k1_group4), modelled
step for step on the CPU: (frame, row, column) advanced by 1 024 pixels per group without divisions, the per-pixel walk of
the slow path, and the two ways a frame's last pixel is recognised.  Narrow and short images make every wrap frequent."""
import pytest


def walk(n, H, W, groups, grid, tids):
    WH, tile = W * H, 1024 * groups
    n_tiles = n * WH // tile
    per = (n_tiles + grid - 1) // grid if n_tiles else 0
    q, r = 1024 // W, 1024 - (1024 // W) * W
    checked = 0
    for b in range(grid):
        t0, t1 = b * per, min(b * per + per, n_tiles)
        for tid in tids:
            if t0 >= t1:
                continue
            px0 = t0 * tile + tid * 4
            f0, r0 = divmod(px0, WH)
            v0, u0 = divmod(r0, W)
            for t in range(t0, t1):
                for g in range(groups):
                    p = t * tile + g * 1024 + tid * 4
                    assert (f0, v0, u0) == (p // WH, (p % WH) // W, p % W)
                    u, v, f, last = u0, v0, f0, -1                  # k1_group4, per-pixel path
                    for j in range(4):
                        assert (f, v, u) == ((p + j) // WH, ((p + j) % WH) // W, (p + j) % W)
                        u += 1
                        if u == W:
                            u, v = 0, v + 1
                            if v == H:
                                v, f, last = 0, f + 1, j
                    if u0 + 3 < W:                                  # k1_group4, vector path
                        assert (3 if (u0 + 4 == W and v0 + 1 == H) else -1) == last
                    assert (v0 + 1 == H and W - u0 <= 4) == (last >= 0)   # frame end seen by a skipped (all-invalid) warp / tile
                    u0, v0 = u0 + r, v0 + q
                    if u0 >= W:
                        u0, v0 = u0 - W, v0 + 1
                    if v0 >= H:
                        v0, f0 = v0 - H, f0 + 1
                    checked += 1
    return checked


@pytest.mark.parametrize("shape", [(6, 8, 256), (5, 9, 257), (3, 40, 300), (3, 12, 1000), (2, 8, 4097), (40, 8, 258), (3, 375, 1242), (3, 8, 65535),
                                   (9, 8, 511), (4, 33, 1025), (7, 8, 259), (2, 1080, 1920)])
@pytest.mark.parametrize("groups", [1, 2])
def test_group_bookkeeping(shape, groups):
    n, H, W = shape
    assert W >= 256 and H >= 8                                      # what the launcher requires of the bulk kernels
    assert walk(n, H, W, groups, grid=7, tids=(0, 1, 2, 63, 100, 254, 255)) > 0 or n * W * H < 1024 * groups
