#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric: world points/s (depth -> world + PLY record) and OctoMap scans/s, on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c2|c3|c4|c5] [--impl reference]

--config (default c2, the configuration the headline metric is quoted on; `config.workload` names it):
  c2  KITTI-odometry shape: 4 500 x 1242x375 uint16 depth + poses -> world float32 records.  One "step" = one pass of the fused
      back-projection + pose transform over the whole sequence, inputs and outputs resident in HBM (`value`, CUDA events);
      `e2e` = the same call through the C ABI with HOST (pinned) buffers, copies inside the timed region.  The line also
      carries the OctoMap half of the metric on a FIXED workload of 1 024 consecutive scans of the same sequence (0.1 m, 80 m):
      scans/s and the .bt SHA-256, the same workload at every N, so the SHA is the same at 1 / 2 / 4 / 8 GPUs.
  c3  the same sequence -> insertPointCloud for all 4 500 scans at 0.1 m / 80 m -> .bt (value = scans/s).
  c4  AirSim drone shape: 1 000 x 640x480 uint16 disparity (PSMNet style, d = raw/256), Z = f B / d, octree at 0.05 m.
  c5  10 000 x 1920x1080 frames, frame-sharded: under torchrun rank r owns a contiguous frame range and its slice of the
      merged cloud at a fixed offset; on one GPU the sequence is streamed in chunks.  Octree at 0.1 m over every frame.
Under torchrun (one rank per GPU) the points path has no data-path collective; the OctoMap path exchanges 136-byte
brick-delta records once per round of scans and applies them in global scan order (sharding.py).

`--impl reference` times the CPU restatement of the reference path (oracle/points_oracle.py, numpy float64: the reference
itself is pure Python at ~38 k points/s/core and cannot travel to the GPU box) on all host cores, same config keys.
"""
import argparse
import hashlib
import importlib
import json
import multiprocessing as mp
import os
import shutil
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DEPTH_SCALE = 1.0 / 256.0
BYTES_PER_PX = 2 + 12          # uint16 sample in, xyz float32 out
METRIC = "world points/s (depth->world+PLY record)"
OCTO_METRIC = "OctoMap scans/s (insertPointCloud, max range 80 m)"
SEED = 20261018

CONFIGS = {
    "c2": dict(id=2, W=1242, H=375, frames=4500, intr="KITTI_INTRINSICS", mode="depth", res=0.1, maxrange=80.0, pose_step=0.8, baseline=0.0,
               workload="C2: KITTI-odometry-shape 4500 x 1242x375 uint16 depth + Colmap-style poses -> world float32 xyz (binary PLY body)"),
    "c3": dict(id=2, W=1242, H=375, frames=4500, intr="KITTI_INTRINSICS", mode="depth", res=0.1, maxrange=80.0, pose_step=0.8, baseline=0.0,
               workload="C3: the C2 sequence (4500 x 1242x375) -> OctoMap insertPointCloud at 0.1 m, max range 80 m, every frame one scan -> .bt"),
    "c4": dict(id=4, W=640, H=480, frames=1000, intr="AIRSIM_INTRINSICS", mode="disparity", res=0.05, maxrange=80.0, pose_step=0.2, baseline=0.25,
               workload="C4: AirSim drone shape, 1000 x 640x480 uint16 disparity (PSMNet style, d = raw/256), Z = f B / d, octree at 0.05 m"),
    "c5": dict(id=5, W=1920, H=1080, frames=10000, intr=(1050.0, 1050.0, 959.5, 539.5), mode="depth", res=0.1, maxrange=80.0, pose_step=0.8, baseline=0.0,
               workload="C5: 10000 x 1920x1080 uint16 depth, frame-sharded across the GPUs -> merged world cloud (per-rank offsets) + 0.1 m octree"),
}


def cfg_of(name):
    from oracle import points_oracle as po
    c = dict(CONFIGS[name])
    c["name"] = name
    if isinstance(c["intr"], str):
        c["intr"] = getattr(po, c["intr"])
    c["fB"] = c["intr"][0] * c["baseline"]
    c["kmode"] = po.MODE_DISPARITY if c["mode"] == "disparity" else po.MODE_DEPTH
    return c


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def base_frame_u16(cfg, k):
    """Base frame k (of 32) of a configuration on the host: analytic street scene (oracle generator); disparity configs store
    round(256 f B / Z), 0 where there is no return."""
    from oracle import points_oracle as po
    z16 = po.synth_depth_u16(cfg["W"], cfg["H"], cfg["intr"], SEED + cfg["id"] + k, cfg.get("depth_kind", "street"))
    if cfg["mode"] != "disparity":
        return z16
    z = z16.astype(np.float64) / 256.0
    return np.where(z > 0, np.round(256.0 * cfg["fB"] / np.maximum(z, 1e-9)), 0).clip(0, 65535).astype(np.uint16)


# ------------------------------------------------------------------------------------------------ CPU arm
def _cpu_worker(job):
    from oracle import points_oracle as po
    name, k0, nf = job
    cfg = cfg_of(name)
    t_acc, pts = 0.0, 0
    for k in range(k0, k0 + nf):
        raw = base_frame_u16(cfg, k % 32)
        q, t = po.synth_pose(k, cfg["frames"], cfg["pose_step"])
        t0 = time.perf_counter()
        rinv = po.scipy_transfer(q)
        _, world = po.depth_to_world(raw, cfg["intr"], rinv, t, cfg["kmode"], DEPTH_SCALE, cfg["fB"])
        rec = world.astype(np.float32)          # the PLY record
        t_acc += time.perf_counter() - t0
        pts += rec.shape[0]
    return pts, t_acc


def cpu_reference_pass(cfg, frames_per_core, cores, pool):
    """One bounded sample of the workload on `cores` processes.  Returns (points, seconds) where seconds is the slowest
    worker's compute time (pose inverse + back-projection + transform + float32 cast; synthesising the frames is not counted)."""
    jobs = [(cfg["name"], c * frames_per_core, frames_per_core) for c in range(cores)]
    res = pool.map(_cpu_worker, jobs)
    return sum(r[0] for r in res), max(r[1] for r in res)


def cpu_baseline_sample(cfg, min_seconds=5.0):
    cores = os.cpu_count() or 1
    with mp.get_context("fork").Pool(cores) as pool:
        cpu_reference_pass(cfg, 1, cores, pool)
        fpc = 2
        pts, sec = cpu_reference_pass(cfg, fpc, cores, pool)
        while sec < min_seconds and fpc < 64:
            fpc *= 2
            pts, sec = cpu_reference_pass(cfg, fpc, cores, pool)
    return {"value": pts / sec, "unit": "points/s", "cores": cores, "kind": "port",
            "sample": "%d frames (%d per core) of the %s sequence, numpy float64 oracle port, %.1f s" % (fpc * cores, fpc, cfg["name"].upper(), sec)}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cfg = cfg_of(args.config)
    cores = os.cpu_count() or 1
    frames_per_core = 4 if cfg["W"] * cfg["H"] < 10 ** 6 else 1
    with mp.get_context("fork").Pool(cores) as pool:
        for _ in range(args.warmup):
            cpu_reference_pass(cfg, 1, cores, pool)
        pts, wall = 0, 0.0
        for _ in range(args.steps):
            p, sec = cpu_reference_pass(cfg, frames_per_core, cores, pool)
            pts += p
            wall += sec
    value = pts / wall
    sample = "%d frames/step (%d per core) of the %s sequence, synthetic street scene, numpy float64 oracle port" % (frames_per_core * cores, frames_per_core, cfg["name"].upper())
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "points/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": cfg["workload"], "frames_per_gpu": cfg["frames"], "frame_range_of_rank0": [0, cfg["frames"]],
                   "pixels_per_step_per_gpu": cfg["frames"] * cfg["W"] * cfg["H"], "frames_per_step": frames_per_core * cores,
                   "frames_per_launch": None, "l2": None, "parity_sample_ok": None, "merged_cloud": None,
                   "sample": sample, "note": "same workload and keys as the GPU arm; a step of THIS arm is a bounded sample of it (frames_per_step frames: the "
                                             "full 4 500-frame pass takes minutes on the host), the value is a rate"},
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


# ------------------------------------------------------------------------------------------------ synthetic sequences on the device
class Sequence:
    """Synthetic frames of a configuration: 32 analytic base frames made on the host (oracle generator, the data the CPU arm
    uses), expanded on the device to any frame index (offset per frame), and the Colmap-style pose of every frame."""

    def __init__(self, torch, cfg, dev):
        self.torch, self.cfg, self.dev = torch, cfg, dev
        base = np.stack([base_frame_u16(cfg, k) for k in range(32)])
        self.base = torch.from_numpy(base.astype(np.int32)).to(dev).reshape(32, cfg["H"] * cfg["W"])

    def frames(self, first, n, out=None):
        """Frames [first, first + n) as an (n, H, W) int16 tensor holding the uint16 bit patterns."""
        torch, H, W = self.torch, self.cfg["H"], self.cfg["W"]
        if out is None:
            out = torch.empty((n, H, W), dtype=torch.int16, device=self.dev)
        o = out.reshape(-1, H * W)
        step = 128
        for a in range(0, n, step):
            k = torch.arange(first + a, first + min(a + step, n), device=self.dev)
            fr = self.base[k % 32]
            fr = torch.where(fr > 0, (fr + ((k * 37) % 64)[:, None]).clamp_(1, 65535), fr)
            fr = torch.where(fr >= 32768, fr - 65536, fr)          # store the uint16 bit pattern in int16
            o[a:a + fr.shape[0]] = fr.to(torch.int16)
        return out.reshape(-1, H, W)[:n]

    def poses(self, first, n):
        from oracle import points_oracle as po
        ps = [po.synth_pose(k, self.cfg["frames"], self.cfg["pose_step"]) for k in range(first, first + n)]
        return np.stack([p[0] for p in ps]), np.stack([p[1] for p in ps])


def centres(rt_host):
    """Sensor origins of the frames (world position of the camera centre), one per pose-table row."""
    mapping = importlib.import_module("3d_reconstruction_system_b200.mapping")
    return mapping.camera_centres(rt_host)


def k1_call(ctx, cfg, depth, rt, out, n, counts, **kw):
    return ctx.backproject(depth, cfg["intr"], rt=rt, mode=cfg["kmode"], depth_scale=DEPTH_SCALE, fB=cfg["fB"], out=out, shape=(n, cfg["H"], cfg["W"]),
                           counts=counts, **kw)


# ------------------------------------------------------------------------------------------------ OctoMap sections
def octomap_cpu_baseline(world_pts, origin, maxrange, res):
    """OctoMap scans/s on one host core: the C restatement (oracle/octomap_oracle.c) inserting one scan."""
    from oracle import octomap_oracle as oo
    tree = oo.OcTree(res)
    t0 = time.perf_counter()
    tree.insertPointCloud_f32(world_pts, origin, maxrange)
    sec = time.perf_counter() - t0
    return {"value": 1.0 / sec, "unit": "scans/s", "cores": 1, "kind": "port",
            "sample": "1 scan of %d rays, C restatement of insertPointCloud (key sets + tree update), %.1f s" % (world_pts.shape[0], sec)}, tree


def octomap_fixed_workload(args, torch, dist, ctx, dev, cfg, seq, rank, world, with_cpu, S, f0):
    """The OctoMap half of the metric on a FIXED workload: scans [f0, f0 + S) of the sequence (K1's world points of those
    frames, device resident), insertPointCloud at cfg.res / cfg.maxrange, in scan order.  One GPU: one pipelined library call
    (r3d_tree_insert_scans); N GPUs: the scans are dealt out in rounds, every rank ray-casts its share, the brick-delta
    records of a round are exchanged and applied in global scan order with the map partitioned by brick owner, and the
    pieces are gathered at the end (inside the timed region) -- every rank then holds the tree of the 1-GPU run, so
    `bt_sha256` is the same line for every N.  Strong scaling: S does not change with N."""
    octomap = importlib.import_module("3d_reconstruction_system_b200.octomap")
    sharding = importlib.import_module("3d_reconstruction_system_b200.sharding")
    H, W = cfg["H"], cfg["W"]
    res, maxrange = cfg["res"], cfg["maxrange"]
    if world == 1:
        mine = list(range(S))
    else:
        mine = sorted(s for _, parts in sharding.scan_rounds(S, world, args.octomap_scans_per_round) for r, a, n in parts if r == rank
                      for s in range(a, a + n))
    # world points of this rank's scans (K1), device resident; consecutive local slots for consecutive scans
    pts = torch.empty((max(len(mine), 1) * H * W, 3), dtype=torch.float32, device=dev)
    rt_all = np.zeros((S, 12))
    slot = {}
    runs = []                                  # maximal runs of consecutive scan indices
    for s_idx in mine:
        if runs and runs[-1][1] == s_idx:
            runs[-1][1] += 1
        else:
            runs.append([s_idx, s_idx + 1])
    j = 0
    for a, b in runs:
        q, t = seq.poses(f0 + a, b - a)
        rt = ctx.pose_to_rt(q, t)
        rt_all[a:b] = rt
        depth = seq.frames(f0 + a, b - a)
        k1_call(ctx, cfg, depth, torch.from_numpy(rt).to(dev), pts[j * H * W:(j + b - a) * H * W], b - a, np.zeros(b - a, np.uint64))
        for s_idx in range(a, b):
            slot[s_idx] = j
            j += 1
    origins = {s_idx: o for s_idx, o in zip(mine, centres(rt_all[mine]))} if mine else {}
    stream = torch.cuda.ExternalStream(ctx.stream(), device=dev)
    out = {"metric": OCTO_METRIC + " @%g m" % res, "unit": "scans/s", "scans": S, "n_gpus": world, "scaling": "strong",
           "workload": "%d consecutive %s scans (%dx%d rays each), frames [%d, %d) of the sequence, the same at every N" % (S, cfg["name"].upper(), W, H, f0, f0 + S)}
    if world == 1:
        org = np.stack([origins[s] for s in range(S)])
        k3_ms = []

        def run(tree, count):
            tree.insertPointClouds(pts[:count * H * W], org[:count], maxrange=maxrange)
            st = tree.lastScanStats()
            k3_ms.append(ctx.last_kernel_ms())
            return st["steps"], st["rays"]

        warm = octomap.OcTree(res, ctx=ctx)
        run(warm, min(16, S))
        del warm
        ms_runs, host_runs, launches_run, growth, wall_runs = [], [], 0, None, []
        for _ in range(3):                       # a pass is short and has one host turnaround per batch of scans: median of three
            tree = octomap.OcTree(res, ctx=ctx)
            tree.reserve(args.reserve_bricks or (1 << 20))   # capacity hint: no pool regrowth inside the timed region
            ctx.set_blocking(False)
            ctx.synchronize()
            launches0 = ctx.launch_count()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            del k3_ms[:]
            tw0 = time.perf_counter()
            steps, rays = run(tree, S)
            tw1 = time.perf_counter()
            e1.record(stream)
            ctx.synchronize()
            tw2 = time.perf_counter()
            ctx.set_blocking(True)
            ms_runs.append(e0.elapsed_time(e1))
            wall_runs.append([1e3 * (tw1 - tw0), 1e3 * (tw2 - tw1)])
            host_runs.append(tree.pipelineStats())
            launches_run = ctx.launch_count() - launches0
            growth = tree.growthStats()
        ms = float(np.median(ms_runs))
        out.update({"value": S / (ms * 1e-3), "ms_per_scan": ms / S, "ms_per_scan_runs": [m / S for m in ms_runs], "rays_per_scan": rays // S,
                    "dda_steps_per_scan": steps // S, "rays_per_s": rays / (ms * 1e-3), "dda_steps_per_s": steps / (ms * 1e-3),
                    "gpu_launches": launches_run, "host_pipeline_runs": host_runs, "host_wall_ms_runs_call_and_drain": wall_runs, "growth": growth,
                    "raycast_kernel_ms_per_scan_last_batch": float(k3_ms[-1]),
                    "raycast_steps_per_s_in_kernel": (steps / S) / max(k3_ms[-1], 1e-9) * 1e3,
                    "timing": "CUDA events on the context stream around one r3d_tree_insert_scans call, median of three passes into fresh trees"})
    else:
        def get_scan(s_idx):
            k = slot[s_idx]
            return pts[k * H * W:(k + 1) * H * W], origins[s_idx]

        def get_scan_batch(first, n):
            k = slot[first]
            return pts[k * H * W:(k + n) * H * W], [H * W] * n, np.stack([origins[first + i] for i in range(n)])

        shared = {}

        def run_once():
            tree = octomap.OcTree(res, ctx=ctx)
            tree.reserve(args.reserve_bricks or (1 << 20))
            sh = sharding.OctreeSharder(tree, get_scan, maxrange=maxrange, owner_partition=True, rank=rank, world=world,
                                        get_scan_batch=get_scan_batch, state=shared)
            ctx.synchronize()
            torch.cuda.synchronize()
            dist.barrier()
            t0 = time.perf_counter()
            sh.run(S, scans_per_rank=args.octomap_scans_per_round)
            ctx.synchronize()
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            sharding.gather_bricks(tree, state=shared)
            ctx.synchronize()
            torch.cuda.synchronize()
            dist.barrier()
            t2 = time.perf_counter()
            return tree, t2 - t0, t2 - t1

        run_once()                                   # warm-up (allocations, NCCL channels)
        secs, merges = [], []
        for _ in range(3):
            tree, sec, merge_sec = run_once()
            tm = torch.tensor([sec, merge_sec], device=dev, dtype=torch.float64)
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            secs.append(float(tm[0].item()))
            merges.append(float(tm[1].item()))
        sec = float(np.median(secs))
        out.update({"value": S / sec, "ms_per_scan": 1e3 * sec / S, "ms_per_scan_runs": [1e3 * x / S for x in secs], "brick_gather_s_runs": merges,
                    "scans_per_round_per_rank": args.octomap_scans_per_round, "exchange": shared.get("exchange"),
                    "timing": "wall clock between barriers + device synchronize, max over ranks, brick gather INSIDE the timed region, median of three passes into fresh trees"})
    t0 = time.perf_counter()
    bt = tree.writeBinary()
    out.update({"voxels": tree.numVoxels(), "bricks": tree.numBricks(), "bt_bytes": len(bt), "bt_write_s": time.perf_counter() - t0,
                "bt_sha256": hashlib.sha256(bt).hexdigest()})
    if world > 1:
        digs = [None] * world
        dist.all_gather_object(digs, out["bt_sha256"])
        out["bt_identical_on_all_ranks"] = len(set(digs)) == 1
    if world == 1 and cfg["name"] == "c2":
        # the mode the reference's own OctoMap scripts use: updateNode(point, True) per point (octomap/txt_transfer_octomap.py:25)
        un = octomap.OcTree(res, ctx=ctx)
        un.reserve(1 << 17)
        n_un = min(S, 12) * H * W                    # 5.6 M points at the KITTI shape: the size ply_transfer_octomap.py caps at (5.4 M)
        un_runs = []
        for _ in range(5):
            un.updateNodes(pts[:H * W], True)
            un.clear()
            ctx.synchronize()
            e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e2.record(stream)
            un.updateNodes(pts[:n_un], True)
            e3.record(stream)
            ctx.synchronize()
            un_runs.append(e2.elapsed_time(e3))
        un_ms = float(np.median(un_runs))
        out["update_node"] = {"metric": "updateNode(point, True) points/s (the reference scripts' mode)", "value": n_un / (un_ms * 1e-3), "unit": "points/s",
                              "points": n_un, "ms": un_ms, "ms_runs": un_runs, "voxels": un.numVoxels()}
    if world == 1 and with_cpu:
        from oracle import octomap_oracle as oo
        w0 = pts[:H * W].cpu().numpy()
        out["cpu_baseline"], ref = octomap_cpu_baseline(w0, origins[0], maxrange, res)
        chk = octomap.OcTree(res, ctx=ctx)            # parity of what was timed: first scan's tree against the oracle
        chk.insertPointCloud(pts[:H * W], origins[0], maxrange=maxrange)
        out["parity_bt_ok"] = bool(chk.writeBinary() == ref.write_binary_bytes())
        if "update_node" in out:
            ref2, chk2 = oo.OcTree(res), octomap.OcTree(res, ctx=ctx)
            t0 = time.perf_counter()
            ref2.updateNodes_f32(w0, True)
            sec = time.perf_counter() - t0
            chk2.updateNodes(pts[:H * W], True)
            out["update_node"]["cpu_baseline"] = {"value": w0.shape[0] / sec, "unit": "points/s", "cores": 1, "kind": "port",
                                                  "sample": "%d points, C restatement of updateNode, %.2f s" % (w0.shape[0], sec)}
            out["update_node"]["parity_bt_ok"] = bool(chk2.writeBinary() == ref2.write_binary_bytes())
    return out


def octomap_full_sequence(args, torch, ctx, dev, cfg, seq, first, n_frames, chunk):
    """insertPointCloud for EVERY frame of [first, first + n_frames) in frame order (configs 3-5) on ONE GPU: the sequence is
    streamed in chunks -- K1 back-projects a chunk into a device buffer, one pipelined library call inserts its scans -- into
    one tree that grows as it goes (growth events reported)."""
    octomap = importlib.import_module("3d_reconstruction_system_b200.octomap")
    H, W = cfg["H"], cfg["W"]
    stream = torch.cuda.ExternalStream(ctx.stream(), device=dev)
    pts = torch.empty((chunk * H * W, 3), dtype=torch.float32, device=dev)
    depth = torch.empty((chunk, H, W), dtype=torch.int16, device=dev)
    # warm the pipeline on a scratch tree (buffers shaped, kernels loaded)
    warm = octomap.OcTree(cfg["res"], ctx=ctx)
    nw = min(8, n_frames)
    q, t = seq.poses(first, nw)
    rt = ctx.pose_to_rt(q, t)
    seq.frames(first, nw, depth)
    k1_call(ctx, cfg, depth[:nw], torch.from_numpy(rt).to(dev), pts[:nw * H * W], nw, np.zeros(nw, np.uint64))
    warm.insertPointClouds(pts[:nw * H * W], centres(rt), maxrange=cfg["maxrange"])
    del warm
    tree = octomap.OcTree(cfg["res"], ctx=ctx)
    if args.reserve_bricks:
        tree.reserve(args.reserve_bricks)
    ins_ms, k1_ms, steps, rays, host = 0.0, 0.0, 0, 0, {"wait_ms": 0.0, "work_ms": 0.0, "max_turnaround_ms": 0.0}
    wall0 = time.perf_counter()
    for a in range(0, n_frames, chunk):
        n = min(chunk, n_frames - a)
        q, t = seq.poses(first + a, n)
        rt = ctx.pose_to_rt(q, t)
        seq.frames(first + a, n, depth)
        rt_dev = torch.from_numpy(rt).to(dev)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ctx.set_blocking(False)
        ev[0].record(stream)
        k1_call(ctx, cfg, depth[:n], rt_dev, pts[:n * H * W], n, np.zeros(n, np.uint64))
        ev[1].record(stream)
        tree.insertPointClouds(pts[:n * H * W], centres(rt), maxrange=cfg["maxrange"])
        ev[2].record(stream)
        ctx.synchronize()
        ctx.set_blocking(True)
        k1_ms += ev[0].elapsed_time(ev[1])
        ins_ms += ev[1].elapsed_time(ev[2])
        st, ps = tree.lastScanStats(), tree.pipelineStats()
        steps += st["steps"]
        rays += st["rays"]
        host["wait_ms"] += ps["wait_ms"]
        host["work_ms"] += ps["work_ms"]
        host["max_turnaround_ms"] = max(host["max_turnaround_ms"], ps["max_turnaround_ms"])
    wall = time.perf_counter() - wall0
    t0 = time.perf_counter()
    bt = tree.writeBinary()
    bt_s = time.perf_counter() - t0
    return {"metric": OCTO_METRIC + " @%g m" % cfg["res"], "value": n_frames / (ins_ms * 1e-3), "unit": "scans/s", "scans": n_frames, "n_gpus": 1,
            "ms_per_scan": ins_ms / n_frames, "insert_ms_total": ins_ms, "k1_ms_total": k1_ms, "wall_s_with_synthesis": wall,
            "rays_per_scan": rays // max(n_frames, 1), "dda_steps_per_scan": steps // max(n_frames, 1), "dda_steps_per_s": steps / (ins_ms * 1e-3),
            "rays_per_s": rays / (ins_ms * 1e-3), "host_pipeline": host, "growth": tree.growthStats(), "voxels": tree.numVoxels(),
            "bricks": tree.numBricks(), "bt_bytes": len(bt), "bt_write_s": bt_s, "bt_sha256": hashlib.sha256(bt).hexdigest(),
            "chunk_frames": chunk, "raycast_kernel_ms_per_scan_last_batch": ctx.last_kernel_ms(),
            "timing": "CUDA events on the context stream around every r3d_tree_insert_scans call (one per chunk of frames), summed; K1 and the synthesis of the frames are outside",
            "workload": "every frame of %s [%d, %d) one scan, in frame order, one tree" % (cfg["name"].upper(), first, first + n_frames)}


# ------------------------------------------------------------------------------------------------ secondary sections (c2)
def compaction_section(torch, ctx, dev, cfg, depth, rt, n=1024):
    """Secondary K1 figure: the same frames with invalid pixels (Z = 0 sky) compacted away, order kept (device resident)."""
    H, W = cfg["H"], cfg["W"]
    n = min(n, depth.shape[0])
    out = torch.empty((n * H * W, 3), dtype=torch.float32, device=dev)
    cnt = torch.zeros(n, dtype=torch.int64, device=dev)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=dev)

    def step():
        k1_call(ctx, cfg, depth[:n], rt[:n], out, n, cnt, compact=True)

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
    peak, _ = load_peaks()
    alg = px * 2 + valid * 12
    return {"frames": n, "ms": ms, "input_pixels_per_s": px / (ms * 1e-3), "valid_fraction": valid / px,
            "algorithmic_gbs": alg / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": alg / (ms * 1e-3) / 1e9 / peak}


def text_section(torch, ctx, dev, cfg, depth, rt, n=64, with_cpu=True):
    """K6 beside K1: the ASCII the reference scripts actually write.  World points of n frames (one streaming
    batch of the drop-in scripts: 64 frames; float64, device resident) -> genply's "%.4f %.4f %.4f \\n" rows and the txt files' str(float64) rows, device to device.
    Timed with CUDA events on the context stream (each call contains one 8-byte size read-back)."""
    import ctypes as C
    H, W = cfg["H"], cfg["W"]
    n = min(n, depth.shape[0])
    npts = n * H * W
    pts = torch.empty((npts, 3), dtype=torch.float64, device=dev)
    k1_call(ctx, cfg, depth[:n], rt[:n], pts, n, np.zeros(n, np.uint64))
    cap = npts * 80
    buf = torch.empty(cap, dtype=torch.uint8, device=dev)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=dev)
    p0 = pts.data_ptr()
    need = C.c_size_t(0)
    peak, _ = load_peaks()
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
        alg = npts * 24 + need.value          # coordinates read once, text written once
        out[name] = {"ms": m, "points_per_s": npts / (m * 1e-3), "text_bytes": int(need.value), "text_gbs": need.value / (m * 1e-3) / 1e9,
                     "algorithmic_gbs": alg / (m * 1e-3) / 1e9, "frac_of_hbm_peak": alg / (m * 1e-3) / 1e9 / peak}
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


def png_decode_section(cfg, n_frames=64):
    """a1 beside the GPU numbers: the frame decode that feeds the path.  The reference reads one PNG at a time with
    cv.imread on one core (transfer/camera_to_world.py:160); the native decoder inflates a batch on every core."""
    import shutil
    import tempfile
    formats = importlib.import_module("3d_reconstruction_system_b200.formats")
    try:
        import cv2
    except Exception:
        return None
    d = tempfile.mkdtemp(prefix="r3d_png_")
    try:
        paths = []
        for k in range(n_frames):
            img = base_frame_u16(cfg, k % 8)
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
                "pixels_per_s": n_frames * cfg["W"] * cfg["H"] / t_native}
    finally:
        shutil.rmtree(d, ignore_errors=True)


def script_e2e_section(cfg, n_frames=256, root=None):
    """Script-level end to end (what the reference IS: files in -> files out, transfer/camera_to_world.py:138-174): n PNG
    depth files + a pose file -> the drop-in script's streamed driver -> one ASCII PLY, wall clock, with the reference's
    per-point text pipeline beside it on a bounded sample.  (The per-frame side files are off: they are 45 MB per frame.)"""
    import shutil
    import tempfile
    try:
        import cv2
    except Exception:
        return None
    from oracle import points_oracle as po
    transfer = importlib.import_module("3d_reconstruction_system_b200.transfer")
    formats = importlib.import_module("3d_reconstruction_system_b200.formats")
    d = tempfile.mkdtemp(prefix="r3d_e2e_", dir=root)
    cwd = os.getcwd()
    try:
        os.chdir(d)
        for sub in ("depth", "point", "point_world", "ply", "camera_pose"):
            os.mkdir(sub)
        names, qs, ts = [], [], []
        for k in range(n_frames):
            img = base_frame_u16(cfg, k % 8)
            names.append("%06d.png" % k)
            cv2.imwrite(os.path.join("depth", names[-1]), np.roll(img, k, axis=1))
            q, t = po.synth_pose(k, cfg["frames"], cfg["pose_step"])
            qs.append(q)
            ts.append(t)
        formats.write_pose_file("camera_pose/image_colmap_simi_2.txt", np.stack(ts), np.stack(qs), names)
        in_bytes = sum(os.path.getsize(os.path.join("depth", nm)) for nm in names)
        transfer.get_file_name("camera_pose/image_colmap_simi_2.txt", write_intermediate=False, ply_path="ply/warm.ply", max_frames=8, quiet=True)
        t0 = time.perf_counter()
        transfer.get_file_name("camera_pose/image_colmap_simi_2.txt", write_intermediate=False, ply_path="ply/out.ply", quiet=True, keep_points=False)
        sec = time.perf_counter() - t0
        out_bytes = os.path.getsize("ply/out.ply")
        npts = n_frames * cfg["W"] * cfg["H"]
        sha = hashlib.sha256()
        with open("ply/out.ply", "rb") as f:
            for blk in iter(lambda: f.read(1 << 24), b""):
                sha.update(blk)
        # the reference's per-point cost on the same frames: its text pipeline on a bounded sample (one core, as it runs)
        k = 100000
        img0 = cv2.imread(os.path.join("depth", names[0]), cv2.IMREAD_GRAYSCALE)
        t0 = time.perf_counter()
        X, Y, Z = po.backproject(po.raw_to_z(img0), po.REF_INTRINSICS)
        cam_txt = po.txt_lines_camera(X.ravel()[:k], Y.ravel()[:k], img0.ravel()[:k])
        rinv = po.scipy_transfer(qs[0])
        wpts = np.array([[float(v) for v in ln.split(',')] for ln in cam_txt.splitlines()])
        w = po.point_camera(wpts, rinv, ts[0])
        _ = po.txt_lines_world(w)
        _ = "".join("%.4f %.4f %.4f \n" % (a, b, c) for a, b, c in w)
        ref_sec = time.perf_counter() - t0
        return {"frames": n_frames, "points": npts, "seconds": sec, "points_per_s": npts / sec, "frames_per_s": n_frames / sec, "png_bytes_in": in_bytes,
                "files_under": os.path.dirname(d),
                "ply_bytes_out": out_bytes, "ply_sha256": sha.hexdigest(),
                "path": "PNG files -> native batch decode (host threads, next batch decoded while this one runs) -> pinned stack -> K1 (fp64 world points stay on the GPU) -> K6 PLY rows -> text read back -> appended to the PLY",
                "cpu_baseline": {"value": k / ref_sec, "unit": "points/s", "cores": 1, "kind": "port",
                                 "sample": "%d pixels through the reference's per-point text pipeline (camera txt -> parse -> transform -> world txt -> PLY row), numpy for the arithmetic" % k}}
    finally:
        os.chdir(cwd)
        shutil.rmtree(d, ignore_errors=True)


# ------------------------------------------------------------------------------------------------ points path
def points_pass(args, torch, dist, ctx, dev, cfg, seq, first, n_frames, chunk):
    """The fused back-projection + transform over frames [first, first + n_frames) of this rank, device resident, in chunks of
    `chunk` frames (one chunk = one launch; the whole range is one chunk when it fits the GPU).  Returns timing + parity."""
    from oracle import points_oracle as po
    H, W = cfg["H"], cfg["W"]
    stream = torch.cuda.ExternalStream(ctx.stream(), device=dev)
    n_chunks = (n_frames + chunk - 1) // chunk
    out = torch.empty((chunk * H * W, 3), dtype=torch.float32, device=dev)
    depth = torch.empty((chunk, H, W), dtype=torch.int16, device=dev)
    st = {"chunk": -1}

    def load(c):
        a = c * chunk
        n = min(chunk, n_frames - a)
        if st["chunk"] != c:
            q, t = seq.poses(first + a, n)
            rt_host = ctx.pose_to_rt(q, t)
            seq.frames(first + a, n, depth)
            st.update(chunk=c, n=n, rt_host=rt_host, rt=torch.from_numpy(rt_host).to(dev), counts=np.zeros(n, np.uint64))
            torch.cuda.synchronize()

    def step_chunk():
        k1_call(ctx, cfg, depth[:st["n"]], st["rt"], out[:st["n"] * H * W], st["n"], st["counts"])

    ctx.set_blocking(False)
    load(0)
    for _ in range(max(args.warmup, 3)):
        step_chunk()
    ctx.synchronize()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(dev.index, getattr(torch.cuda.get_device_properties(dev), "uuid", None))
    sampler.start()
    launches0 = ctx.launch_count()
    evs = []
    for i in range(args.steps):
        for c in range(n_chunks):
            load(c)                         # (synthesising a chunk's frames is outside the events)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            step_chunk()
            e1.record(stream)
            evs.append((e0, e1))
            if n_chunks > 1:
                ctx.synchronize()
    ctx.synchronize()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    clocks = sampler.stop()
    launches = ctx.launch_count() - launches0
    chunk_ms = [a.elapsed_time(b) for a, b in evs]
    per_step = [sum(chunk_ms[i * n_chunks:(i + 1) * n_chunks]) for i in range(args.steps)]
    # one resident chunk: the launches run back to back, first event to last event (what round 1 reported); streamed: sum
    total_ms = evs[0][0].elapsed_time(evs[-1][1]) if n_chunks == 1 else sum(per_step)
    if dist is not None:
        tm = torch.tensor([total_ms], device=dev)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        total_ms = float(tm.item())
    ctx.set_blocking(True)
    # sampled parity check of what was just timed (never a fallback: it only asserts)
    k = st["n"] // 2
    raw = depth[k].cpu().numpy().view(np.uint16)
    ref = po.depth_to_world(raw, cfg["intr"], st["rt_host"][k, :9].reshape(3, 3), st["rt_host"][k, 9:], cfg["kmode"], DEPTH_SCALE, cfg["fB"])[1].astype(np.float32)
    got = out[k * H * W:(k + 1) * H * W].cpu().numpy()
    st["ref_frame"] = (k, ref)
    return {"total_ms_max": total_ms, "per_step_ms": per_step, "launches": int(launches), "clocks": clocks, "parity_ok": bool(np.array_equal(got, ref)),
            "depth": depth, "out": out, "state": st, "n_chunks": n_chunks}


def pcie_ceiling(world):
    """Bare pinned D2H rate of this pool's boxes with `world` ranks copying at once (tools/pcie_probe.py under torchrun,
    committed as profiles/r2_pcie_probe.json)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r2_pcie_probe.json"))).get(str(world))
    except Exception:
        return None


def e2e_pass(args, torch, dist, ctx, dev, cfg, depth, st, world):
    """The same call through the C ABI with pinned HOST buffers: uploads, kernels and read-backs inside the timed region."""
    import ctypes as C
    H, W = cfg["H"], cfg["W"]
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = 64 << 30
    frames = st["n"]
    while frames > 16 and frames * H * W * 14 * 1.3 > avail * 0.5 / max(world, 1):   # every rank of the node pins its own buffers
        frames //= 2
    lib = ctx.lib
    in_bytes, out_bytes = frames * H * W * 2, frames * H * W * 12
    h_in, h_out = lib.r3d_host_alloc(in_bytes), lib.r3d_host_alloc(out_bytes)
    e2e = None
    if h_in and h_out:
        np_in = np.ctypeslib.as_array(C.cast(h_in, C.POINTER(C.c_uint16)), shape=(frames, H, W))
        np_in[:] = depth[:frames].cpu().numpy().view(np.uint16)
        np_out = np.ctypeslib.as_array(C.cast(h_out, C.POINTER(C.c_float)), shape=(frames * H * W, 3))
        counts = np.zeros(frames, np.uint64)
        e_steps = max(1, min(args.steps, 3))

        def call():
            ctx.backproject(np_in, cfg["intr"], rt=st["rt_host"][:frames], mode=cfg["kmode"], depth_scale=DEPTH_SCALE, fB=cfg["fB"], out=np_out, counts=counts)

        call()
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            call()
        wall = time.perf_counter() - t0
        if dist is not None:
            tw = torch.tensor([wall], device=dev, dtype=torch.float64)
            dist.all_reduce(tw, op=dist.ReduceOp.MAX)
            wall = float(tw.item())
        k, ref = st["ref_frame"]
        ok = bool(np.array_equal(np_out[k * H * W:(k + 1) * H * W], ref)) if k < frames else None
        d2h_gbs = world * out_bytes * e_steps / wall / 1e9
        e2e = {"value": world * frames * H * W * e_steps / wall, "unit": "points/s", "h2d_bytes_per_step": in_bytes + frames * 96,
               "d2h_bytes_per_step": out_bytes, "frames": frames, "steps": e_steps, "parity_ok": ok, "d2h_gbs_aggregate": d2h_gbs,
               "path": "r3d_backproject_rt with pinned host buffers (ring of three device slots: uploads, kernels and read-backs on three streams)"}
        ceil = pcie_ceiling(world)
        if ceil and ceil.get("d2h_gbs_aggregate"):
            e2e["pcie_ceiling"] = ceil
            e2e["frac_of_measured_d2h_ceiling"] = d2h_gbs / ceil["d2h_gbs_aggregate"]
        del np_in, np_out
    lib.r3d_host_free(h_in)
    lib.r3d_host_free(h_out)
    return e2e


# ------------------------------------------------------------------------------------------------ GPU arm
def run_gpu_arm(args):
    import torch
    r3d = importlib.import_module("3d_reconstruction_system_b200")
    sharding = importlib.import_module("3d_reconstruction_system_b200.sharding")
    cfg = cfg_of(args.config)
    cfg["depth_kind"] = args.depth_kind
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # CPU baseline first (rank 0, N=1 only): worker processes are forked before any CUDA state exists
    with_cpu = world == 1 and rank == 0 and not args.no_cpu_baseline
    cpu = cpu_baseline_sample(cfg) if with_cpu else None
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    ctx = r3d.Context(local_rank)
    if args.frames:
        cfg["frames"] = args.frames
    total_frames = cfg["frames"]
    seq = Sequence(torch, cfg, dev)
    H, W = cfg["H"], cfg["W"]
    if cfg["name"] == "c5":
        # frame-sharded: rank r owns a contiguous range; its slice of the merged cloud starts at a fixed offset (strong scaling)
        lo, hi = sharding.frame_range(total_frames, world, rank)
        scaling = "strong"
    else:
        # the per-GPU sequence is fixed (weak scaling): N independent shards of the same shape
        lo, hi = 0, total_frames
        scaling = "weak"
    n_frames = hi - lo
    free_b, _ = torch.cuda.mem_get_info(dev)
    chunk = max(1, min(n_frames, int(min(free_b * 0.55, 80e9) // (H * W * 14))))
    pp = points_pass(args, torch, dist, ctx, dev, cfg, seq, lo, n_frames, chunk)
    px_rank = n_frames * H * W
    px_total = px_rank * world if scaling == "weak" else total_frames * H * W
    value = px_total * args.steps / (pp["total_ms_max"] * 1e-3)
    e2e = e2e_pass(args, torch, dist, ctx, dev, cfg, pp["depth"], pp["state"], world)

    sections = {}
    if cfg["name"] == "c2" and world == 1 and not args.quick:
        sections["compact_mode"] = compaction_section(torch, ctx, dev, cfg, pp["depth"], pp["state"]["rt"])
        try:
            sections["text_rows"] = text_section(torch, ctx, dev, cfg, pp["depth"], pp["state"]["rt"], with_cpu=with_cpu)
        except Exception as exc:                      # a secondary figure must not take the headline line down
            sections["text_rows"] = {"error": str(exc)[:200]}
    pp.pop("depth")
    pp.pop("out")
    torch.cuda.empty_cache()

    octo = None
    if cfg["name"] in ("c2", "c4") and args.octomap_scans > 0:
        S = min(args.octomap_scans, total_frames)
        octo = octomap_fixed_workload(args, torch, dist, ctx, dev, cfg, seq, rank, world, with_cpu, S, total_frames // 2 - S // 2)
    elif cfg["name"] in ("c3", "c5") and not args.no_octomap:
        S = min(args.octomap_scans_full or total_frames, total_frames)
        if world == 1:
            ochunk = max(8, min(256, int(12e9 // (H * W * 14))))
            octo = octomap_full_sequence(args, torch, ctx, dev, cfg, seq, 0, S, ochunk)
            if with_cpu:
                q, t = seq.poses(total_frames // 2, 1)
                rt1 = ctx.pose_to_rt(q, t)
                pts1 = torch.empty((H * W, 3), dtype=torch.float32, device=dev)
                k1_call(ctx, cfg, seq.frames(total_frames // 2, 1), torch.from_numpy(rt1).to(dev), pts1, 1, np.zeros(1, np.uint64))
                octo["cpu_baseline"], _ = octomap_cpu_baseline(pts1.cpu().numpy(), centres(rt1)[0], cfg["maxrange"], cfg["res"])
        else:
            octo = octomap_fixed_workload(args, torch, dist, ctx, dev, cfg, seq, rank, world, False, S, 0)
    png = script = None
    if rank == 0 and with_cpu and cfg["name"] == "c2" and not args.quick:
        png = png_decode_section(cfg)
        try:
            script = script_e2e_section(cfg)
            # the same run with the files on tmpfs: what the box's disk (3.25 GB of PLY through the page cache) costs the figure
            shm = "/dev/shm"
            try:
                if script and os.path.isdir(shm) and shutil.disk_usage(shm).free > (8 << 30):
                    mem = script_e2e_section(cfg, root=shm)
                    script["on_tmpfs"] = {k: mem[k] for k in ("seconds", "points_per_s", "frames_per_s", "files_under", "ply_sha256")}
            except Exception as exc:
                script["on_tmpfs"] = {"error": str(exc)[:300]}
        except Exception as exc:
            script = {"error": str(exc)[:300]}
    if rank == 0:
        peak, peak_src = load_peaks()
        kernel_ms = float(np.mean(pp["per_step_ms"])) / pp["n_chunks"]
        px_launch = min(chunk, n_frames) * H * W
        achieved = px_launch * BYTES_PER_PX / (kernel_ms * 1e-3) / 1e9
        traffic = None
        if cfg["name"] in ("c2", "c3") and n_frames == 4500:
            try:
                traffic = json.load(open(os.path.join(ROOT, "profiles", "k1_traffic.json"))).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        headline_octo = cfg["name"] == "c3" and octo is not None
        line = {
            "metric": octo["metric"] if headline_octo else METRIC, "value": octo["value"] if headline_octo else value,
            "unit": "scans/s" if headline_octo else "points/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": (octo["ms_per_scan"] * octo["scans"]) if headline_octo else pp["total_ms_max"] / args.steps, "higher_is_better": True,
            "scaling": octo["scaling"] if (headline_octo and "scaling" in octo) else scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": cfg["workload"] + ("" if total_frames == CONFIGS[cfg["name"]]["frames"] else " [--frames %d]" % total_frames),
                       "frames_per_gpu": n_frames, "frame_range_of_rank0": [lo, hi], "pixels_per_step_per_gpu": px_rank, "frames_per_step": n_frames,
                       "frames_per_launch": min(chunk, n_frames),
                       "l2": "inputs+outputs %.1f GB per launch >> 126 MB L2" % (px_launch * 14 / 1e9), "parity_sample_ok": pp["parity_ok"],
                       "merged_cloud": ("rank r writes its records at offset r * %d points of the merged cloud; no data-path collective" % px_rank) if world > 1 else None},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": peak_src, "kernel": "k1_bulk_vec<u16,f32,world,%s>" % cfg["mode"], "bytes_per_pixel": BYTES_PER_PX,
                         "kernel_ms": kernel_ms, "pixels_per_launch": px_launch},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": pp["launches"], "clocks": pp["clocks"], "points": {"value": value, "unit": "points/s"},
            "octomap": octo, "png_decode": png, "script_e2e": script,
        }
        line.update(sections)
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
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS), help="BASELINE.json configuration (default: c2, the headline)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="skip the secondary sections (compaction, text rows, PNG decode, script e2e)")
    ap.add_argument("--octomap-scans", type=int, default=1024, help="scans of the fixed OctoMap workload of c2 / c4 (0 = skip)")
    ap.add_argument("--octomap-scans-full", type=int, default=0, help="c3 / c5: scans to insert (default: every frame)")
    ap.add_argument("--no-octomap", action="store_true", help="c3 / c5: skip the octree")
    ap.add_argument("--octomap-scans-per-round", type=int, default=32, help="multi-GPU: scans each rank ray-casts between two exchanges")
    ap.add_argument("--reserve-bricks", type=int, default=0, help="capacity hint for the map (c3 / c5 default: let it grow, growth events are reported)")
    ap.add_argument("--depth-kind", default="street", choices=["street", "uniform"],
                    help="synthetic depth: analytic street scene (headline) or i.i.d. U[1, 80] m (ray-casting worst case, SURVEY.md section 8d)")
    ap.add_argument("--frames", type=int, default=0, help="frames of the sequence (default: the configuration's)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
