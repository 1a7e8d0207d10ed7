#!/usr/bin/env python
"""bench.py -- headline metric of BASELINE.json: world points/s for depth -> world (+PLY record).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--frames F]

One "step" = one pass of the hot path (fused back-projection + pose transform, float32 PLY records) over the
KITTI-odometry-shape sequence of BASELINE config 2: 4 500 synthetic 1242x375 uint16 depth frames with poses.
`value` times the kernel with inputs and outputs resident in HBM (CUDA events on the launching stream);
`e2e` times the same call through the C ABI with HOST (pinned) buffers, H2D and D2H copies inside the timed region.
Under torchrun every rank runs its own shard of frames (weak scaling: the per-GPU sequence is fixed), no data-path
collective; the time is the max over ranks.

`--impl reference` times the CPU restatement of the reference path (oracle/points_oracle.py, numpy float64: the
reference itself is pure Python at ~38 k points/s/core and cannot travel to the GPU box) on all host cores.
"""
import argparse
import importlib
import json
import multiprocessing as mp
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H = 1242, 375
N_FRAMES_C2 = 4500
DEPTH_SCALE = 1.0 / 256.0
BYTES_PER_PX = 2 + 12          # uint16 depth in, xyz float32 out
METRIC = "world points/s (depth->world+PLY record)"
WORKLOAD = "C2: KITTI-odometry-shape 4500 x 1242x375 uint16 depth + Colmap-style poses -> world float32 xyz (binary PLY body)"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------ CPU arm
def _cpu_worker(args):
    from oracle import points_oracle as po
    seed, k0, nf, n_total = args
    intr = po.KITTI_INTRINSICS
    t_acc = 0.0
    pts = 0
    for k in range(k0, k0 + nf):
        depth = po.synth_depth_u16(W, H, intr, seed + (k % 32), "street")
        q, t = po.synth_pose(k, n_total)
        t0 = time.perf_counter()
        rinv = po.scipy_transfer(q)
        _, world = po.depth_to_world(depth, intr, rinv, t, po.MODE_DEPTH, DEPTH_SCALE)
        rec = world.astype(np.float32)          # the PLY record
        t_acc += time.perf_counter() - t0
        pts += rec.shape[0]
    return pts, t_acc


def cpu_reference_pass(frames_per_core, cores, pool):
    """One bounded sample of the workload on `cores` processes.  Returns (points, seconds) where seconds is the
    slowest worker's compute time (pose inverse + back-projection + transform + float32 cast; synthesising the
    depth frames is not counted)."""
    jobs = [(20261018 + 2, c * frames_per_core, frames_per_core, N_FRAMES_C2) for c in range(cores)]
    res = pool.map(_cpu_worker, jobs)
    return sum(r[0] for r in res), max(r[1] for r in res)


def cpu_baseline_sample(min_seconds=5.0):
    cores = os.cpu_count() or 1
    with mp.get_context("fork").Pool(cores) as pool:
        cpu_reference_pass(1, cores, pool)
        fpc = 2
        pts, sec = cpu_reference_pass(fpc, cores, pool)
        while sec < min_seconds and fpc < 64:
            fpc *= 2
            pts, sec = cpu_reference_pass(fpc, cores, pool)
    return {"value": pts / sec, "unit": "points/s", "cores": cores, "kind": "port",
            "sample": "%d frames (%d per core) of the C2 sequence, numpy float64 oracle port, %.1f s" % (fpc * cores, fpc, sec)}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    frames_per_core = 4
    with mp.get_context("fork").Pool(cores) as pool:
        for _ in range(args.warmup):
            cpu_reference_pass(1, cores, pool)
        pts, wall = 0, 0.0
        for _ in range(args.steps):
            p, sec = cpu_reference_pass(frames_per_core, cores, pool)
            pts += p
            wall += sec
    value = pts / wall
    sample = "%d frames/step (%d per core) of the C2 sequence, synthetic street depth, numpy float64 oracle port" % (frames_per_core * cores, frames_per_core)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "points/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "points/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    def __init__(self, index, uuid=None):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = None
            if uuid:
                try:
                    self.h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(uuid)).encode())
                except Exception:
                    self.h = None
            if self.h is None:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.02)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------ GPU arm
def synth_on_device(torch, n_frames, rank, dev, kind="street"):
    """Synthetic depth of the named shape: 32 analytic 'street' frames made on the host (oracle generator, the same
    data the CPU arm uses), expanded on the device to n_frames distinct frames (shift + offset per frame)."""
    from oracle import points_oracle as po
    base = np.stack([po.synth_depth_u16(W, H, po.KITTI_INTRINSICS, 20261018 + 2 + k, kind) for k in range(32)])
    b = torch.from_numpy(base.astype(np.int32)).to(dev).reshape(32, H * W)
    out = torch.empty((n_frames, H * W), dtype=torch.int16, device=dev)
    step = 256
    for a in range(0, n_frames, step):
        k = torch.arange(a, min(a + step, n_frames), device=dev)
        fr = b[k % 32]
        fr = torch.where(fr > 0, (fr + ((k * 37) % 64)[:, None]).clamp_(1, 65535), fr)
        fr = torch.where(fr >= 32768, fr - 65536, fr)          # store the uint16 bit pattern in int16
        out[a:a + fr.shape[0]] = fr.to(torch.int16)
    poses = [po.synth_pose(k, max(n_frames, 1)) for k in range(n_frames)]
    q = np.stack([p[0] for p in poses])
    t = np.stack([p[1] for p in poses])
    return out.reshape(n_frames, H, W), q, t


def octomap_cpu_baseline(world_pts, origin, maxrange, res):
    """OctoMap scans/s on one host core: the C restatement (oracle/octomap_oracle.c) inserting one scan."""
    from oracle import octomap_oracle as oo
    tree = oo.OcTree(res)
    t0 = time.perf_counter()
    tree.insertPointCloud_f32(world_pts, origin, maxrange)
    sec = time.perf_counter() - t0
    return {"value": 1.0 / sec, "unit": "scans/s", "cores": 1, "kind": "port",
            "sample": "1 scan of %d rays, C restatement of insertPointCloud (key sets + tree update), %.1f s" % (world_pts.shape[0], sec)}


def octomap_section(args, torch, r3d, ctx, dev, depth, rt_host, n_scans, with_cpu):
    """Second half of BASELINE.json's metric: OctoMap scans/s @0.1 m, max range 80 m (config 3).  Scans are the
    world points of consecutive frames of the same sequence (K1 output, device resident); each scan is ray-cast
    into a brick delta and applied in order.  Timed with CUDA events on the context stream."""
    from oracle import points_oracle as po
    octomap = importlib.import_module("3d_reconstruction_system_b200.octomap")
    n_scans = min(n_scans, depth.shape[0])
    f0 = depth.shape[0] // 2 - n_scans // 2          # middle of the trajectory: every key in range
    world = torch.empty((n_scans * H * W, 3), dtype=torch.float32, device=dev)
    rt = torch.from_numpy(rt_host[f0:f0 + n_scans].copy()).to(dev)
    ctx.backproject(depth[f0:f0 + n_scans], po.KITTI_INTRINSICS, rt=rt, depth_scale=DEPTH_SCALE, out=world, shape=(n_scans, H, W),
                    counts=np.zeros(n_scans, np.uint64))
    origins = [po.camera_centre(rt_host[f0 + i, :9].reshape(3, 3), rt_host[f0 + i, 9:]) for i in range(n_scans)]
    stream = torch.cuda.ExternalStream(ctx.stream(), device=dev)
    res, maxrange = 0.1, 80.0

    k3_ms = []

    def run(tree, count):
        # one library call for the whole batch of scans (r3d_tree_insert_scans): no interpreter time between scans
        tree.insertPointClouds(world[:count * H * W], origins[:count], maxrange=maxrange)
        st = tree.lastScanStats()
        k3_ms.append(ctx.last_kernel_ms())          # ray-cast kernel of the last scan of the batch
        return st["steps"], st["rays"]

    warm = octomap.OcTree(res, ctx=ctx)
    run(warm, min(8, n_scans))
    del warm
    # the timed batch, three times on a fresh tree (median): one pass is ~30 ms and the scan pipeline has a host turnaround
    # per scan, so a single pass is at the mercy of one descheduled host thread
    ms_runs, host_runs, launches_run = [], [], 0
    for _ in range(3):
        tree = octomap.OcTree(res, ctx=ctx)
        tree.reserve(1 << 17)            # capacity hint (277 MB): no pool regrowth inside the timed region
        ctx.set_blocking(False)          # scans are queued back to back; the events below bracket the device work
        ctx.synchronize()
        launches0 = ctx.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        del k3_ms[:]
        steps, rays = run(tree, n_scans)
        e1.record(stream)
        ctx.synchronize()
        ctx.set_blocking(True)
        ms_runs.append(e0.elapsed_time(e1))
        host_runs.append(tree.pipelineStats())
        launches_run = ctx.launch_count() - launches0
    ms = float(np.median(ms_runs))
    t0 = time.perf_counter()
    bt = tree.writeBinary()
    bt_s = time.perf_counter() - t0
    out = {"metric": "OctoMap scans/s @0.1 m (insertPointCloud, max range 80 m)", "value": n_scans / (ms * 1e-3), "unit": "scans/s",
           "scans": n_scans, "rays_per_scan": rays // n_scans, "dda_steps_per_scan": steps // n_scans,
           "rays_per_s": rays / (ms * 1e-3), "dda_steps_per_s": steps / (ms * 1e-3), "ms_per_scan": ms / n_scans,
           "voxels": tree.numVoxels(), "bricks": tree.numBricks(), "bt_bytes": len(bt), "bt_write_s": bt_s,
           "gpu_launches": launches_run, "ms_per_scan_runs": [m / n_scans for m in ms_runs], "host_pipeline_runs": host_runs,
           "raycast_kernel_ms_last_scan": float(k3_ms[-1]), "raycast_steps_per_s_in_kernel": (steps / n_scans) / max(k3_ms[-1], 1e-9) * 1e3,
           "workload": "C3: %d consecutive KITTI-shape %s scans (1242x375 rays each, Z=0 sky pixels included as rays to the sensor origin)" % (n_scans, args.depth_kind)}
    # the mode the reference's own OctoMap scripts use: updateNode(point, True) per point (octomap/txt_transfer_octomap.py:25)
    un = octomap.OcTree(res, ctx=ctx)
    un.reserve(1 << 17)
    n_un = min(n_scans, 12) * H * W                      # 5.6 M points: the size ply_transfer_octomap.py caps at (5.4 M)
    un_runs = []
    for _ in range(5):                                   # median of 5: a ~1.4 ms call with two counter read-backs
        un.updateNodes(world[:H * W], True)
        un.clear()
        ctx.synchronize()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record(stream)
        un.updateNodes(world[:n_un], True)
        e3.record(stream)
        ctx.synchronize()
        un_runs.append(e2.elapsed_time(e3))
    un_ms = float(np.median(un_runs))
    out["update_node"] = {"metric": "updateNode(point, True) points/s (the reference scripts' mode)", "value": n_un / (un_ms * 1e-3), "unit": "points/s",
                          "points": n_un, "ms": un_ms, "ms_runs": un_runs, "voxels": un.numVoxels(),
                          "note": "median of 5 passes, each into a cleared tree; the first pass also allocates the staging buffers for this batch size"}
    if with_cpu:
        w0 = world[:H * W].cpu().numpy()
        out["cpu_baseline"] = octomap_cpu_baseline(w0, origins[0], maxrange, res)
        # parity of what was timed: first scan's tree against the oracle
        from oracle import octomap_oracle as oo
        chk, ref = octomap.OcTree(res, ctx=ctx), oo.OcTree(res)
        chk.insertPointCloud(world[:H * W], origins[0], maxrange=maxrange)
        ref.insertPointCloud_f32(w0, origins[0], maxrange)
        out["parity_bt_ok"] = bool(chk.writeBinary() == ref.write_binary_bytes())
        # updateNode mode: CPU oracle on a bounded sample (one frame's points), and .bt parity of that sample
        ref2, chk2 = oo.OcTree(res), octomap.OcTree(res, ctx=ctx)
        t0 = time.perf_counter()
        ref2.updateNodes_f32(w0, True)
        sec = time.perf_counter() - t0
        chk2.updateNodes(world[:H * W], True)
        out["update_node"]["cpu_baseline"] = {"value": w0.shape[0] / sec, "unit": "points/s", "cores": 1, "kind": "port",
                                              "sample": "%d points, C restatement of updateNode, %.2f s" % (w0.shape[0], sec)}
        out["update_node"]["parity_bt_ok"] = bool(chk2.writeBinary() == ref2.write_binary_bytes())
    return out


def compaction_section(torch, ctx, dev, depth, rt, n=1024):
    """Secondary K1 figure: the same frames with invalid pixels (Z = 0 sky) compacted away, order kept (device resident)."""
    from oracle import points_oracle as po
    n = min(n, depth.shape[0])
    out = torch.empty((n * H * W, 3), dtype=torch.float32, device=dev)
    cnt = torch.zeros(n, dtype=torch.int64, device=dev)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=dev)

    def step():
        ctx.backproject(depth[:n], po.KITTI_INTRINSICS, rt=rt[:n], depth_scale=DEPTH_SCALE, out=out, shape=(n, H, W), counts=cnt, compact=True)

    for _ in range(3):
        step()
    ctx.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(5):
        step()
    e1.record(stream)
    ctx.synchronize()
    ms = e0.elapsed_time(e1) / 5
    valid = int(cnt.sum().item())
    px = n * H * W
    return {"frames": n, "ms": ms, "input_pixels_per_s": px / (ms * 1e-3), "valid_fraction": valid / px,
            "algorithmic_gbs": (px * 2 + valid * 12) / (ms * 1e-3) / 1e9,
            "kernels": "k1_count_tiles (warp per tile) + CUB scan + k1_bulk_compact (validity nibbles + warp scan, in-tile packing in shared memory, 16-byte stores for all-valid warps, empty tiles / warps skipped, bulk stores)"}


def text_section(torch, ctx, dev, depth, rt, n=16, with_cpu=True):
    """K6 beside K1: the ASCII the reference scripts actually write.  World points of n frames (float64,
    device resident) -> genply's "%.4f %.4f %.4f \n" rows and the txt files' str(float64) rows, device to device.
    Timed with CUDA events on the context stream (each call contains one 8-byte size read-back)."""
    import ctypes as C
    from oracle import points_oracle as po
    n = min(n, depth.shape[0])
    npts = n * H * W
    pts = torch.empty((npts, 3), dtype=torch.float64, device=dev)
    ctx.backproject(depth[:n], po.KITTI_INTRINSICS, rt=rt[:n], depth_scale=DEPTH_SCALE, out=pts, shape=(n, H, W), counts=np.zeros(n, np.uint64))
    cap = npts * 80
    buf = torch.empty(cap, dtype=torch.uint8, device=dev)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=dev)
    p0 = pts.data_ptr()
    need = C.c_size_t(0)
    out = {"points": npts}
    for name in ("ply", "txt"):
        def call():
            if name == "ply":
                rc = ctx.lib.r3d_format_ply_rows(ctx.handle, p0, p0 + 8, p0 + 16, 3, npts, None, buf.data_ptr(), cap, C.byref(need))
            else:
                rc = ctx.lib.r3d_format_txt_rows(ctx.handle, p0, p0 + 8, p0 + 16, 3, npts, 0, buf.data_ptr(), cap, C.byref(need))
            if rc != 0:
                raise RuntimeError("K6 %s rows failed: rc %d" % (name, rc))
        call()
        ctx.synchronize()
        ms = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            call()
            e1.record(stream)
            ctx.synchronize()
            ms.append(e0.elapsed_time(e1))
        m = float(np.median(ms))
        out[name] = {"ms": m, "points_per_s": npts / (m * 1e-3), "text_bytes": int(need.value), "text_gbs": need.value / (m * 1e-3) / 1e9,
                     "algorithmic_gbs": (npts * 24 * 2 + need.value) / (m * 1e-3) / 1e9}
        if with_cpu:
            k = 200000                                   # the reference's own loop (camera_to_world.py:117-121 / :103), one core
            h = pts[:k].cpu().numpy()
            t0 = time.perf_counter()
            if name == "ply":
                ref = "".join("%.4f %.4f %.4f \n" % (a, b, c) for a, b, c in h).encode()
            else:
                ref = "".join(str(a) + "," + str(b) + "," + str(c) + "\n" for a, b, c in h).encode()
            sec = time.perf_counter() - t0
            got = bytes(buf[:len(ref)].cpu().numpy())
            out[name]["cpu_baseline"] = {"value": k / sec, "unit": "points/s", "cores": 1, "kind": "reference",
                                         "sample": "%d points through the reference's Python formatting expression" % k}
            out[name]["parity_ok"] = bool(got == ref)
    return out


def png_decode_section(n_frames=64):
    """a1 beside the GPU numbers: the frame decode that feeds the path.  The reference reads one PNG at a time with
    cv.imread on one core (transfer/camera_to_world.py:160); the native decoder inflates a batch on every core."""
    import shutil
    import tempfile
    from oracle import points_oracle as po
    formats = importlib.import_module("3d_reconstruction_system_b200.formats")
    try:
        import cv2
    except Exception:
        return None
    d = tempfile.mkdtemp(prefix="r3d_png_")
    try:
        paths = []
        for k in range(n_frames):
            img = po.synth_depth_u16(W, H, po.KITTI_INTRINSICS, 20261018 + 2 + (k % 8), "street")
            p = os.path.join(d, "%05d.png" % k)
            cv2.imwrite(p, np.roll(img, k, axis=1))
            paths.append(p)
        size = sum(os.path.getsize(p) for p in paths)
        formats.imread_batch(paths[:8], "raw")
        t0 = time.perf_counter()
        got = formats.imread_batch(paths, "raw")
        t_native = time.perf_counter() - t0
        t0 = time.perf_counter()
        ref = [cv2.imread(p, cv2.IMREAD_UNCHANGED) for p in paths[:16]]
        t_cv = (time.perf_counter() - t0) * n_frames / 16
        ok = all(np.array_equal(got[i], ref[i]) for i in range(16))
        return {"frames": n_frames, "png_bytes_per_frame": size // n_frames, "native_frames_per_s": n_frames / t_native, "threads": os.cpu_count(),
                "cv2_imread_frames_per_s_1core": n_frames / t_cv, "identical_to_cv2": bool(ok),
                "pixels_per_s": n_frames * W * H / t_native}
    finally:
        shutil.rmtree(d, ignore_errors=True)


def octomap_section_multi(args, torch, dist, r3d, ctx, dev, depth, rt_host, scans_per_gpu, rank, world):
    """OctoMap scans/s on N GPUs: S = scans_per_gpu * N consecutive scans of the sequence; every rank
    ray-casts only its share, the brick-delta records of each round are all-gathered over NCCL and applied in global scan
    order with the map partitioned by brick owner; at the end the per-rank pieces are gathered, so every rank holds the same
    tree as a 1-GPU run.  Timed by wall clock between barriers (the exchange runs on torch's NCCL stream), max over ranks."""
    from oracle import points_oracle as po
    octomap = importlib.import_module("3d_reconstruction_system_b200.octomap")
    sharding = importlib.import_module("3d_reconstruction_system_b200.sharding")
    S = scans_per_gpu * world
    S = min(S, depth.shape[0])
    f0 = depth.shape[0] // 2 - S // 2
    res, maxrange = 0.1, 80.0
    mine = sorted(s for _, parts in sharding.scan_rounds(S, world, args.octomap_scans_per_round) for r, a, n in parts if r == rank
                  for s in range(a, a + n))
    pts = torch.empty((max(len(mine), 1) * H * W, 3), dtype=torch.float32, device=dev)
    slot = {}
    for j, s_idx in enumerate(mine):        # world points of this rank's scans (K1), device resident
        k = f0 + s_idx
        ctx.backproject(depth[k:k + 1], po.KITTI_INTRINSICS, rt=torch.from_numpy(rt_host[k:k + 1].copy()).to(dev), depth_scale=DEPTH_SCALE,
                        out=pts[j * H * W:(j + 1) * H * W], shape=(1, H, W), counts=np.zeros(1, np.uint64))
        slot[s_idx] = j
    origins = {s_idx: po.camera_centre(rt_host[f0 + s_idx, :9].reshape(3, 3), rt_host[f0 + s_idx, 9:]) for s_idx in mine}

    def get_scan(s_idx):
        j = slot[s_idx]
        return pts[j * H * W:(j + 1) * H * W], origins[s_idx]

    def get_scan_batch(first, n):
        j = slot[first]
        return pts[j * H * W:(j + n) * H * W], [H * W] * n, np.stack([origins[first + i] for i in range(n)])

    shared = {}

    def run_once():
        tree = octomap.OcTree(res, ctx=ctx)
        tree.reserve(1 << 17)
        sh = sharding.OctreeSharder(tree, get_scan, maxrange=maxrange, owner_partition=True, rank=rank, world=world,
                                    get_scan_batch=get_scan_batch)
        sh._buf = shared.get("buf")          # record buffer sized by the warm-up run
        ctx.synchronize()
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        sh.run(S, scans_per_rank=args.octomap_scans_per_round)
        ctx.synchronize()
        torch.cuda.synchronize()
        dist.barrier()
        t1 = time.perf_counter()
        sharding.gather_bricks(tree)
        ctx.synchronize()
        torch.cuda.synchronize()
        dist.barrier()
        t2 = time.perf_counter()
        shared["buf"] = sh._buf
        return tree, t1 - t0, t2 - t1

    run_once()                                   # warm-up (allocations, NCCL channels)
    tree, sec, merge_sec = run_once()
    tm = torch.tensor([sec, merge_sec], device=dev, dtype=torch.float64)
    dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    sec, merge_sec = float(tm[0].item()), float(tm[1].item())
    bt = tree.writeBinary()
    import hashlib
    digest = hashlib.sha256(bt).hexdigest()
    digs = [None] * world
    dist.all_gather_object(digs, digest)
    return {"metric": "OctoMap scans/s @0.1 m (insertPointCloud, max range 80 m)", "value": S / sec, "unit": "scans/s", "scans": S,
            "n_gpus": world, "scaling": "weak", "scans_per_gpu": scans_per_gpu, "ms_per_scan": 1e3 * sec / S, "brick_gather_s": merge_sec,
            "voxels": tree.numVoxels(), "bt_bytes": len(bt), "bt_sha256": digest, "bt_identical_on_all_ranks": len(set(digs)) == 1,
            "exchange": "NCCL all-gather of 136-byte brick-delta records per round of %d scans per rank; owner-partitioned apply (one library call per round and peer); brick gather at the end" % args.octomap_scans_per_round,
            "timing": "wall clock between barriers + device synchronize, max over ranks",
            "workload": "C3: %d consecutive KITTI-shape street scans (1242x375 rays each)" % S}


def run_gpu_arm(args):
    import torch
    r3d = importlib.import_module("3d_reconstruction_system_b200")
    from oracle import points_oracle as po

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # CPU baseline first (rank 0, N=1 only): worker processes are forked before any CUDA state exists
    cpu = cpu_baseline_sample() if (world == 1 and rank == 0 and not args.no_cpu_baseline) else None
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    ctx = r3d.Context(local_rank)
    n_frames = args.frames or N_FRAMES_C2
    depth, q, t = synth_on_device(torch, n_frames, rank, dev, args.depth_kind)
    rt_host = ctx.pose_to_rt(q, t)
    rt = torch.from_numpy(rt_host).to(dev)
    px = n_frames * H * W
    out = torch.empty((px, 3), dtype=torch.float32, device=dev)
    counts = np.zeros(n_frames, np.uint64)
    torch.cuda.synchronize()
    stream = torch.cuda.ExternalStream(ctx.stream(), device=dev)

    def step():
        ctx.backproject(depth, po.KITTI_INTRINSICS, rt=rt, depth_scale=DEPTH_SCALE, out=out, shape=(n_frames, H, W), counts=counts)

    ctx.set_blocking(False)
    for _ in range(max(args.warmup, 3)):
        step()
    ctx.synchronize()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank, getattr(torch.cuda.get_device_properties(dev), "uuid", None))
    sampler.start()
    launches0 = ctx.launch_count()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    evs[0].record(stream)
    for i in range(args.steps):
        step()
        evs[i + 1].record(stream)
    ctx.synchronize()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    clocks = sampler.stop()
    launches = ctx.launch_count() - launches0
    total_ms = evs[0].elapsed_time(evs[-1])
    per_step = [evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)]
    if dist is not None:
        tm = torch.tensor([total_ms], device=dev)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        total_ms_max = float(tm.item())
    else:
        total_ms_max = total_ms
    ctx.set_blocking(True)
    value = world * px * args.steps / (total_ms_max * 1e-3)

    # sampled parity check of what was just timed (never a fallback: it only asserts)
    k = n_frames // 2
    ref = po.depth_to_world(depth[k].cpu().numpy().view(np.uint16), po.KITTI_INTRINSICS, rt_host[k, :9].reshape(3, 3), rt_host[k, 9:],
                            po.MODE_DEPTH, DEPTH_SCALE)[1].astype(np.float32)
    got = out[k * H * W:(k + 1) * H * W].cpu().numpy()
    parity_ok = bool(np.array_equal(got, ref))

    # ---- end to end through the C ABI with host buffers
    e2e = None
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = 64 << 30
    e2e_frames = n_frames
    # every rank of the node pins its own host buffers: share half of the free memory between them
    while e2e_frames > 64 and e2e_frames * H * W * 14 * 1.3 > avail * 0.5 / max(world, 1):
        e2e_frames //= 2
    lib = ctx.lib
    in_bytes, out_bytes = e2e_frames * H * W * 2, e2e_frames * H * W * 12
    h_in, h_out = lib.r3d_host_alloc(in_bytes), lib.r3d_host_alloc(out_bytes)
    if h_in and h_out:
        import ctypes as C
        np_in = np.ctypeslib.as_array(C.cast(h_in, C.POINTER(C.c_uint16)), shape=(e2e_frames, H, W))
        np_in[:] = depth[:e2e_frames].cpu().numpy().view(np.uint16)
        np_out = np.ctypeslib.as_array(C.cast(h_out, C.POINTER(C.c_float)), shape=(e2e_frames * H * W, 3))
        e_steps = max(1, min(args.steps, 3))
        ctx.backproject(np_in, po.KITTI_INTRINSICS, rt=rt_host[:e2e_frames], depth_scale=DEPTH_SCALE, out=np_out, counts=counts)
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            ctx.backproject(np_in, po.KITTI_INTRINSICS, rt=rt_host[:e2e_frames], depth_scale=DEPTH_SCALE, out=np_out, counts=counts)
        wall = time.perf_counter() - t0
        if dist is not None:
            tw = torch.tensor([wall], device=dev, dtype=torch.float64)
            dist.all_reduce(tw, op=dist.ReduceOp.MAX)
            wall = float(tw.item())
        e2e_ok = bool(np.array_equal(np_out[k * H * W:(k + 1) * H * W], ref)) if k < e2e_frames else None
        e2e = {"value": world * e2e_frames * H * W * e_steps / wall, "unit": "points/s", "h2d_bytes_per_step": in_bytes + e2e_frames * 96,
               "d2h_bytes_per_step": out_bytes, "frames": e2e_frames, "steps": e_steps, "parity_ok": e2e_ok,
               "path": "r3d_backproject_rt with pinned host buffers (ring of three device slots: uploads, kernels and read-backs on three streams)"}
        del np_in, np_out
    lib.r3d_host_free(h_in)
    lib.r3d_host_free(h_out)

    compaction = compaction_section(torch, ctx, dev, depth, rt) if world == 1 else None
    text = None
    if world == 1:
        try:
            text = text_section(torch, ctx, dev, depth, rt, with_cpu=not args.no_cpu_baseline)
        except Exception as exc:                      # a secondary figure must not take the headline line down
            text = {"error": str(exc)[:200]}
    octo = None
    if args.octomap_scans > 0:
        del out
        torch.cuda.empty_cache()
        if world == 1:
            octo = octomap_section(args, torch, r3d, ctx, dev, depth, rt_host, args.octomap_scans, with_cpu=not args.no_cpu_baseline)
        else:
            octo = octomap_section_multi(args, torch, dist, r3d, ctx, dev, depth, rt_host, args.octomap_scans, rank, world)

    png = png_decode_section() if (rank == 0 and not args.no_cpu_baseline) else None
    if rank == 0:
        peak, peak_src = load_peaks()
        kernel_ms = float(np.mean(per_step))
        achieved = px * BYTES_PER_PX / (kernel_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "k1_traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": value, "unit": "points/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD if n_frames == N_FRAMES_C2 else WORKLOAD.replace("4500", str(n_frames)),
                       "frames_per_gpu": n_frames, "pixels_per_step_per_gpu": px, "l2": "inputs+outputs %.1f GB per step >> 126 MB L2" % (px * 14 / 1e9),
                       "parity_sample_ok": parity_ok},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": peak_src, "kernel": "k1_bulk_vec<u16,f32,world>", "bytes_per_pixel": BYTES_PER_PX,
                         "kernel_ms": kernel_ms},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "octomap": octo, "png_decode": png, "compact_mode": compaction, "text_rows": text,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    ctx.close()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--octomap-scans", type=int, default=32, help="scans for the OctoMap scans/s section (0 = skip)")
    ap.add_argument("--octomap-scans-per-round", type=int, default=4, help="multi-GPU: scans each rank ray-casts between two exchanges")
    ap.add_argument("--depth-kind", default="street", choices=["street", "uniform"],
                    help="synthetic depth: analytic street scene (headline) or i.i.d. U[1, 80] m (ray-casting worst case, SURVEY.md section 8d)")
    ap.add_argument("--frames", type=int, default=0, help="frames per GPU (default: the 4500 of BASELINE config 2)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
