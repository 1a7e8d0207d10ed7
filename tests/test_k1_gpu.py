"""Parity of the fused back-projection / pose-transform kernel (K1) with the oracle, through the C ABI.

Bars (BASELINE.json north_star): points within 1e-5 relative or 1e-4 m absolute of the float64 numpy reference;
voxel keys bit-exact.  The kernel mirrors the oracle's operation order in fp64, so the tests below assert the much
stronger property that fp64 records are BIT-IDENTICAL to the oracle and fp32 records are its correctly rounded cast."""
import json
import os

import numpy as np
import pytest

from oracle import points_oracle as po

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(r3d):
    c = r3d.Context(0)
    yield c
    c.close()


def oracle_batch(depths, intr, rts, mode=0, depth_scale=1.0, fB=0.0):
    cams, worlds = [], []
    for k in range(depths.shape[0]):
        rinv = rts[k, :9].reshape(3, 3) if rts is not None else np.eye(3)
        t = rts[k, 9:] if rts is not None else np.zeros(3)
        cam, world = po.depth_to_world(depths[k], intr, rinv, t, mode, depth_scale, fB)
        cams.append(cam)
        worlds.append(world)
    return np.concatenate(cams), np.concatenate(worlds)


def random_rt(n, rng, spread=1700.0):
    q = rng.normal(size=(n, 4))
    t = rng.uniform(-spread, spread, size=(n, 3))
    return np.concatenate([np.stack([po.scipy_transfer(q[k]).reshape(9) for k in range(n)]), t], axis=1)


def keys_of(xyz32, res=0.1):
    return np.floor((1.0 / res) * xyz32.astype(np.float64)).astype(np.int64) + 32768


def assert_tolerance(got, ref):
    err = np.abs(got.astype(np.float64) - ref)
    assert np.all((err <= 1e-4) | (err <= 1e-5 * np.abs(ref)))


def test_golden_reference_frames(ctx, golden_dir):
    """The three frames the reference itself processed (tests/golden, made by oracle/gen_golden.py)."""
    z = np.load(os.path.join(golden_dir, "ref_c2w_small.npz"))
    rt = np.concatenate([z["rinv"].reshape(3, 9), z["trans"]], axis=1)
    got64, counts = ctx.backproject(z["depths"], po.REF_INTRINSICS, rt=rt, out_dtype=np.float64)
    ref = z["world"].reshape(-1, 3)
    assert_tolerance(got64, ref)
    scale = np.abs(z["trans"]).max(axis=1).repeat(77)[:, None] + np.abs(ref) + 255.0
    assert np.max(np.abs(got64 - ref) / scale) <= 4 * np.finfo(np.float64).eps      # reference BLAS order: <= 2 ulp
    _, oracle_world = oracle_batch(z["depths"], po.REF_INTRINSICS, rt)
    assert np.array_equal(got64, oracle_world)                                        # fixed-order oracle: bit-exact
    got32, _ = ctx.backproject(z["depths"], po.REF_INTRINSICS, rt=rt)
    assert np.array_equal(got32, oracle_world.astype(np.float32))
    assert np.array_equal(keys_of(got32), keys_of(ref.astype(np.float32)))
    assert counts.tolist() == [77, 77, 77]
    # through the quaternion entry point (C++ pose conversion): still within tolerance and key-identical
    got_q, _ = ctx.backproject_qt(z["depths"], po.REF_INTRINSICS, z["quats"], z["trans"])
    assert_tolerance(got_q, ref)
    assert np.array_equal(keys_of(got_q), keys_of(ref.astype(np.float32)))


def test_camera_frame_matches_reference_text(ctx, golden_dir):
    """gentxtcord: camera-frame fp64 records reproduce the reference's txt bytes."""
    z = np.load(os.path.join(golden_dir, "ref_c2w_small.npz"))
    txt = json.load(open(os.path.join(golden_dir, "ref_c2w_small_text.json")))
    cam, _ = ctx.backproject(z["depths"], po.REF_INTRINSICS, rt=None, out_dtype=np.float64)
    cam = cam.reshape(3, -1, 3)
    for k in range(3):
        assert po.txt_lines_camera(cam[k, :, 0], cam[k, :, 1], z["depths"][k]) == txt["cam_txt"][k]
    g = np.load(os.path.join(golden_dir, "ref_p2c_640.npz"))
    cam, _ = ctx.backproject(g["depth"], po.REF_INTRINSICS, rt=None, out_dtype=np.float64)
    assert np.array_equal(cam[g["sel"], 0], g["X"]) and np.array_equal(cam[g["sel"], 1], g["Y"])
    assert np.array_equal(cam[g["sel"], 2], g["Z"])


@pytest.mark.parametrize("dtype,scale", [(np.uint16, 1.0 / 256.0), (np.uint8, 1.0), (np.float32, 1.0)])
@pytest.mark.parametrize("shape", [(5, 375, 1242), (3, 480, 640), (2, 7, 11), (1, 1, 1), (1, 33, 1025)])
def test_bit_exact_vs_oracle(ctx, dtype, scale, shape):
    rng = np.random.default_rng(20261018 + shape[2])
    n, H, W = shape
    if dtype == np.float32:
        depths = rng.uniform(0, 80, size=shape).astype(np.float32)
    else:
        depths = rng.integers(0, np.iinfo(dtype).max + 1, size=shape).astype(dtype)
    depths.reshape(-1)[0] = 0
    intr = po.KITTI_INTRINSICS if W == 1242 else po.REF_INTRINSICS
    rt = random_rt(n, rng)
    cam_ref, world_ref = oracle_batch(depths, intr, rt, 0, scale)
    got64, _ = ctx.backproject(depths, intr, rt=rt, depth_scale=scale, out_dtype=np.float64)
    assert np.array_equal(got64, world_ref)
    got32, _ = ctx.backproject(depths, intr, rt=rt, depth_scale=scale)
    assert np.array_equal(got32, world_ref.astype(np.float32))
    assert np.array_equal(keys_of(got32), keys_of(world_ref.astype(np.float32)))
    cam64, _ = ctx.backproject(depths, intr, rt=None, depth_scale=scale, out_dtype=np.float64)
    assert np.array_equal(cam64, cam_ref)


def test_disparity_mode(ctx):
    rng = np.random.default_rng(4)
    shape = (3, 480, 640)
    disp = rng.integers(0, 65536, size=shape).astype(np.uint16)
    disp[0, :10] = 0                                       # invalid disparities -> Z = 0
    fB = 269.5 * 0.25
    rt = random_rt(3, rng, spread=100.0)
    _, world_ref = oracle_batch(disp, po.AIRSIM_INTRINSICS, rt, po.MODE_DISPARITY, 1.0 / 256.0, fB)
    got64, _ = ctx.backproject(disp, po.AIRSIM_INTRINSICS, rt=rt, mode=po.MODE_DISPARITY, depth_scale=1.0 / 256.0, fB=fB,
                               out_dtype=np.float64)
    assert np.array_equal(got64, world_ref)


@pytest.mark.parametrize("holes", [0, 1, 2, 3, 5])
def test_compaction_all_valid_warps_every_alignment(ctx, holes):
    """Warps whose 128 pixels are all valid write their records with 16-byte stores at whatever alignment the records
    before them leave: `holes` invalid pixels at the start shift everything after them by 3 * holes words."""
    rng = np.random.default_rng(300 + holes)
    n, H, W = 3, 40, 1026
    depths = rng.integers(1, 65535, size=(n, H, W)).astype(np.uint16)
    depths[0, 0, :holes] = 0
    depths[1, 20, 500] = 0                                 # one more shift in the middle of the batch
    rt = random_rt(n, rng)
    _, world_ref = oracle_batch(depths, po.KITTI_INTRINSICS, rt, 0, 1.0 / 256.0)
    mask = np.concatenate([po.valid_mask(depths[k], 0, 1.0 / 256.0).ravel() for k in range(n)])
    got, counts = ctx.backproject(depths, po.KITTI_INTRINSICS, rt=rt, depth_scale=1.0 / 256.0, compact=True)
    assert counts.tolist() == [H * W - holes, H * W - 1, H * W]
    assert np.array_equal(got, world_ref[mask].astype(np.float32))


def test_compaction_frames_ending_in_empty_tiles(ctx):
    """Tiles without a valid pixel are skipped by the bulk compaction kernel; frames that end inside one still get
    their counts (two empty frames in the middle, an empty one at the start)."""
    rng = np.random.default_rng(41)
    n, H, W = 5, 40, 1026                                  # frames end in the middle of 2048-pixel tiles
    depths = rng.integers(1, 65535, size=(n, H, W)).astype(np.uint16)
    depths[0] = 0
    depths[2] = 0
    depths[3] = 0
    rt = random_rt(n, rng)
    _, world_ref = oracle_batch(depths, po.KITTI_INTRINSICS, rt, 0, 1.0 / 256.0)
    mask = np.concatenate([po.valid_mask(depths[k], 0, 1.0 / 256.0).ravel() for k in range(n)])
    got, counts = ctx.backproject(depths, po.KITTI_INTRINSICS, rt=rt, depth_scale=1.0 / 256.0, compact=True)
    assert counts.tolist() == [0, H * W, 0, 0, H * W]
    assert np.array_equal(got, world_ref[mask].astype(np.float32))
    # a scale outside the range where validity is a compare on the sample: the exact fp64 rule decides (1e-320 * raw
    # is a positive denormal, so every non-zero sample stays valid)
    got2, counts2 = ctx.backproject(depths[:2], po.KITTI_INTRINSICS, rt=rt[:2], depth_scale=1e-320, compact=True, out_dtype=np.float64)
    assert counts2.tolist() == [0, H * W]
    _, wref2 = oracle_batch(depths[:2], po.KITTI_INTRINSICS, rt[:2], 0, 1e-320)
    assert np.array_equal(got2, wref2[mask[:2 * H * W]])


def test_compaction_keeps_order_and_counts(ctx):
    rng = np.random.default_rng(5)
    shape = (4, 375, 1242)
    depths = rng.integers(0, 4, size=shape).astype(np.uint16) * rng.integers(0, 20000, size=shape).astype(np.uint16)
    depths[2] = 0                                          # a frame with no valid pixel
    rt = random_rt(4, rng)
    _, world_ref = oracle_batch(depths, po.KITTI_INTRINSICS, rt, 0, 1.0 / 256.0)
    mask = np.concatenate([po.valid_mask(depths[k], 0, 1.0 / 256.0).ravel() for k in range(4)])
    got, counts = ctx.backproject(depths, po.KITTI_INTRINSICS, rt=rt, depth_scale=1.0 / 256.0, compact=True, out_dtype=np.float64)
    assert counts.tolist() == [int(po.valid_mask(depths[k]).sum()) for k in range(4)]
    assert counts[2] == 0
    assert np.array_equal(got, world_ref[mask])
    got32, _ = ctx.backproject(depths, po.KITTI_INTRINSICS, rt=rt, depth_scale=1.0 / 256.0, compact=True)
    assert np.array_equal(got32, world_ref[mask].astype(np.float32))


def test_pitched_rows_and_empty_batch(ctx):
    rng = np.random.default_rng(6)
    n, H, W = 2, 37, 101
    padded = rng.integers(0, 65536, size=(n, H, W + 27)).astype(np.uint16)
    depths = np.ascontiguousarray(padded[:, :, :W])
    rt = random_rt(n, rng)
    _, world_ref = oracle_batch(depths, po.REF_INTRINSICS, rt, 0, 0.001)
    got, _ = ctx.backproject(padded, po.REF_INTRINSICS, rt=rt, depth_scale=0.001, out_dtype=np.float64, shape=(n, H, W),
                             pitch=(W + 27) * 2)
    assert np.array_equal(got, world_ref)
    got0, c0 = ctx.backproject(np.zeros((0, 4, 4), np.uint8), po.REF_INTRINSICS, rt=np.zeros((0, 12)))
    assert got0.shape == (0, 3)


def test_device_pointers_equal_host_path(ctx):
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(8)
    shape = (9, 375, 1242)
    depths = rng.integers(0, 65536, size=shape).astype(np.uint16)
    rt = random_rt(9, rng)
    host, _ = ctx.backproject(depths, po.KITTI_INTRINSICS, rt=rt, depth_scale=1 / 256.0)
    d_depth = torch.from_numpy(depths.view(np.int16)).cuda()
    d_rt = torch.from_numpy(rt).cuda()
    d_out = torch.empty((9 * 375 * 1242, 3), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    before = ctx.launch_count()
    ctx.backproject(d_depth, po.KITTI_INTRINSICS, rt=d_rt, depth_scale=1 / 256.0, out=d_out, shape=shape,
                    counts=np.zeros(9, np.uint64))
    assert ctx.launch_count() > before
    assert np.array_equal(d_out.cpu().numpy(), host)
    assert ctx.last_kernel_ms() > 0


def test_large_batch_properties(ctx):
    """BASELINE-size frames, 256 of them: chunk invariance + sampled oracle check (the full oracle would take minutes)."""
    rng = np.random.default_rng(9)
    n, H, W = 256, 375, 1242
    depths = rng.integers(0, 65536, size=(n, H, W)).astype(np.uint16)
    q = np.stack([po.synth_pose(k, n)[0] for k in range(n)])
    t = np.stack([po.synth_pose(k, n)[1] for k in range(n)])
    rt = ctx.pose_to_rt(q, t)
    whole, _ = ctx.backproject(depths, po.KITTI_INTRINSICS, rt=rt, depth_scale=1 / 256.0)
    parts = [ctx.backproject(depths[a:a + 37], po.KITTI_INTRINSICS, rt=rt[a:a + 37], depth_scale=1 / 256.0)[0] for a in range(0, n, 37)]
    assert np.array_equal(whole, np.concatenate(parts))
    for k in (0, 101, 255):
        _, wref = po.depth_to_world(depths[k], po.KITTI_INTRINSICS, rt[k, :9].reshape(3, 3), rt[k, 9:], 0, 1 / 256.0)
        assert np.array_equal(whole[k * H * W:(k + 1) * H * W], wref.astype(np.float32))


def test_bad_arguments_raise(ctx, r3d):
    with pytest.raises(r3d.R3DError):
        ctx.backproject(np.zeros((1, 4, 4), np.uint8), po.REF_INTRINSICS, mode=7)
    with pytest.raises(TypeError):
        ctx.backproject(np.zeros((1, 4, 4), np.int64), po.REF_INTRINSICS)
    with pytest.raises(ValueError):
        ctx.pose_to_rt(np.zeros((1, 4)), np.zeros((1, 3)))


def test_point_camera_and_transform_points(ctx, golden_dir):
    z = np.load(os.path.join(golden_dir, "ref_c2w_small.npz"))
    import importlib
    tr = importlib.import_module("3d_reconstruction_system_b200.transfer")
    k = 2
    cam, _ = po.depth_to_world(z["depths"][k], po.REF_INTRINSICS, z["rinv"][k], z["trans"][k])
    got = tr.point_camera(cam, z["rinv"][k], z["trans"][k])
    assert np.array_equal(got, po.point_camera(cam, z["rinv"][k], z["trans"][k]))
    assert tr.point_camera(cam[3], z["rinv"][k], z["trans"][k]).shape == (3, 1)
    r = tr.scipy_transfer(z["quats"][k])
    assert np.max(np.abs(np.asarray(r) - z["rinv"][k])) <= 4 * np.finfo(np.float64).eps
    T = np.eye(4)
    T[:3, :3] = z["rinv"][0]
    T[:3, 3] = [1.0, -2.0, 3.5]
    out = ctx.transform_points(cam, T)
    want = np.stack([((T[i, 0] * cam[:, 0] + T[i, 1] * cam[:, 1]) + T[i, 2] * cam[:, 2]) + T[i, 3] for i in range(3)], axis=1)
    assert np.array_equal(out, want)


def test_c5_shape_1920x1080_bit_exact(ctx):
    """BASELINE config 5 frame shape (1920x1080 RGBD): bit-exact against the oracle, float64 and float32 records."""
    rng = np.random.default_rng(55)
    n, H, W = 2, 1080, 1920
    depths = rng.integers(0, 65536, size=(n, H, W)).astype(np.uint16)
    intr = (1050.0, 1050.0, 959.5, 539.5)
    rt = random_rt(n, rng)
    _, world_ref = oracle_batch(depths, intr, rt, 0, 1.0 / 256.0)
    got64, _ = ctx.backproject(depths, intr, rt=rt, depth_scale=1.0 / 256.0, out_dtype=np.float64)
    assert np.array_equal(got64, world_ref)
    got32, _ = ctx.backproject(depths, intr, rt=rt, depth_scale=1.0 / 256.0)
    assert np.array_equal(got32, world_ref.astype(np.float32))


def test_c2_full_size_properties(ctx):
    """BASELINE config 2 at full size (4 500 x 1242x375 frames, 2.1 G points, 29 GB of traffic), device resident:
    the one-launch result equals a chunked recomputation bit for bit (tile / CTA decomposition invariance), frames at
    sampled positions equal the oracle, and re-running is idempotent."""
    torch = pytest.importorskip("torch")
    free, _ = torch.cuda.mem_get_info()
    if free < 70 << 30:
        pytest.skip("needs ~60 GB of free device memory")
    dev = torch.device("cuda", 0)
    n, H, W = 4500, 375, 1242
    g = torch.Generator(device=dev)
    g.manual_seed(20261018)
    depth = torch.randint(0, 32768, (n, H, W), dtype=torch.int16, device=dev, generator=g)      # uint16 bit patterns
    depth[7] = 0
    q = np.stack([po.synth_pose(k, n)[0] for k in range(n)])
    t = np.stack([po.synth_pose(k, n)[1] for k in range(n)])
    rt_h = ctx.pose_to_rt(q, t)
    rt = torch.from_numpy(rt_h).to(dev)
    whole = torch.empty((n * H * W, 3), dtype=torch.float32, device=dev)
    parts = torch.empty((n * H * W, 3), dtype=torch.float32, device=dev)
    cnt = np.zeros(n, np.uint64)
    ctx.backproject(depth, po.KITTI_INTRINSICS, rt=rt, depth_scale=1 / 256.0, out=whole, shape=(n, H, W), counts=cnt)
    assert int(cnt.sum()) == n * H * W
    step = 613                                          # not a divisor of anything in the tiling
    for a in range(0, n, step):
        b = min(n, a + step)
        ctx.backproject(depth[a:b], po.KITTI_INTRINSICS, rt=rt[a:b], depth_scale=1 / 256.0, out=parts[a * H * W:b * H * W], shape=(b - a, H, W),
                        counts=np.zeros(b - a, np.uint64))
    ctx.synchronize()
    assert torch.equal(whole.view(torch.int32), parts.view(torch.int32))
    for k in (0, 7, 2250, 4499):
        d = depth[k].cpu().numpy().view(np.uint16)
        _, wref = po.depth_to_world(d, po.KITTI_INTRINSICS, rt_h[k, :9].reshape(3, 3), rt_h[k, 9:], 0, 1 / 256.0)
        assert np.array_equal(whole[k * H * W:(k + 1) * H * W].cpu().numpy(), wref.astype(np.float32))
    ctx.backproject(depth, po.KITTI_INTRINSICS, rt=rt, depth_scale=1 / 256.0, out=parts, shape=(n, H, W), counts=cnt)
    ctx.synchronize()
    assert torch.equal(whole.view(torch.int32), parts.view(torch.int32))
    del whole, parts, depth
    torch.cuda.empty_cache()


@pytest.mark.parametrize("dtype,scale,out_dtype,world", [(np.uint16, 1.0 / 256.0, np.float32, True), (np.uint8, 1.0, np.float64, True),
                                                         (np.float32, 1.0, np.float32, False), (np.uint16, 1.0 / 256.0, np.float64, False)])
def test_fast_compaction_paths(ctx, dtype, scale, out_dtype, world):
    """The bulk compaction kernel (full 1024-pixel tiles) + generic tail: order, per-frame counts and values for sparse,
    dense, empty and full frames, every 16-byte phase of the output offset, all sample types."""
    rng = np.random.default_rng(77)
    n, H, W = 6, 33, 1025                                  # 202 950 px: 198 full tiles + a 198-px tail, frames end mid-tile
    if dtype == np.float32:
        depths = rng.uniform(0.5, 80, size=(n, H, W)).astype(np.float32)
    else:
        depths = rng.integers(1, np.iinfo(dtype).max, size=(n, H, W)).astype(dtype)
    keep = rng.random(size=(n, H, W))
    depths[0][keep[0] < 0.9] = 0                           # sparse
    depths[1][keep[1] < 0.01] = 0                          # dense
    depths[2] = 0                                          # empty
    depths[4][keep[4] < 0.5] = 0                           # frame 3 stays full
    depths[5, -1, -1] = 0
    rt = random_rt(n, rng) if world else None
    cam_ref, world_ref = oracle_batch(depths, po.REF_INTRINSICS, rt if world else random_rt(n, rng), 0, scale)
    ref = world_ref if world else cam_ref
    mask = np.concatenate([po.valid_mask(depths[k], 0, scale).ravel() for k in range(n)])
    got, counts = ctx.backproject(depths, po.REF_INTRINSICS, rt=rt, depth_scale=scale, compact=True, out_dtype=out_dtype)
    assert counts.tolist() == [int(po.valid_mask(depths[k], 0, scale).sum()) for k in range(n)]
    assert got.shape[0] == int(mask.sum())
    assert np.array_equal(got, ref[mask].astype(out_dtype))
    # disparity mode through the same kernels
    if dtype == np.uint16 and world:
        got, counts = ctx.backproject(depths, po.AIRSIM_INTRINSICS, rt=rt, mode=po.MODE_DISPARITY, depth_scale=scale, fB=67.375, compact=True,
                                      out_dtype=np.float64)
        _, wref = oracle_batch(depths, po.AIRSIM_INTRINSICS, rt, po.MODE_DISPARITY, scale, 67.375)
        assert np.array_equal(got, wref[mask])


@pytest.mark.parametrize("shape", [(6, 8, 256), (5, 9, 257), (3, 40, 300), (3, 12, 1000), (2, 8, 4097), (40, 8, 258)])
def test_narrow_and_short_images_through_the_bulk_kernels(ctx, shape):
    """The narrowest / shortest images the bulk kernels take (W >= 256, H >= 8): several row wraps per 1 024-pixel group,
    frame ends every few tiles; plain and compaction mode, both record types, two sample types."""
    rng = np.random.default_rng(5 + shape[2])
    n, H, W = shape
    for dtype, scale in ((np.uint16, 1.0 / 256.0), (np.float32, 1.0)):
        depths = (rng.integers(0, 60000, size=shape) * (rng.random(shape) > 0.2)).astype(dtype)
        rt = random_rt(n, rng)
        _, world_ref = oracle_batch(depths, po.REF_INTRINSICS, rt, 0, scale)
        mask = np.concatenate([po.valid_mask(depths[k], 0, scale).ravel() for k in range(n)])
        for od in (np.float32, np.float64):
            got, _ = ctx.backproject(depths, po.REF_INTRINSICS, rt=rt, depth_scale=scale, out_dtype=od)
            assert np.array_equal(got, world_ref.astype(od))
            gc, counts = ctx.backproject(depths, po.REF_INTRINSICS, rt=rt, depth_scale=scale, compact=True, out_dtype=od)
            assert counts.tolist() == [int(po.valid_mask(depths[k], 0, scale).sum()) for k in range(n)]
            assert np.array_equal(gc, world_ref[mask].astype(od))
