"""Generate tests/golden/* by RUNNING THE REFERENCE's own functions (authoring container only).

The reference (/root/reference, read-only) is plain Python for the point path, so it is imported here
with the three shims SURVEY.md section 8c documents (stub matplotlib / mpl_toolkits, np.float = float)
and its functions are executed unmodified on small seeded inputs.  Inputs and outputs are committed
as fixtures so that the parity tests run where /root/reference does not exist (the GPU box).

    python oracle/gen_golden.py            # rewrites tests/golden/ref_*.npz / .json

OctoMap has no fixtures: the `octomap` module is absent from the reference and this image
(parity unpinned, see oracle/octomap_oracle.c).
"""
import hashlib
import importlib.util
import json
import os
import sys
import tempfile
import types

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def load_reference():
    for name in ("matplotlib", "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.mplot3d"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["mpl_toolkits.mplot3d"].Axes3D = object
    if not hasattr(np, "float"):
        np.float = float  # removed in numpy >= 1.24; reference uses it at camera_to_world.py:29
    mods = {}
    for key, rel in (("c2w", "transfer/camera_to_world.py"), ("p2c", "transfer/pixel_to_camera.py")):
        spec = importlib.util.spec_from_file_location("ref_" + key, os.path.join(REF, rel))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        mods[key] = m
    return mods


def sha(s):
    return hashlib.sha256(s if isinstance(s, bytes) else s.encode()).hexdigest()


def main():
    import cv2
    os.makedirs(OUT, exist_ok=True)
    ref = load_reference()
    c2w, p2c = ref["c2w"], ref["p2c"]
    rng = np.random.default_rng(20261018)
    meta = {}

    # ---- (1) camera_to_world.py sequence driver on 3 tiny frames, run through get_file_name ----
    H, W = 7, 11
    n_frames = 3
    depths = rng.integers(0, 256, size=(n_frames, H, W), dtype=np.uint8)
    depths[0, 0, 0] = 0
    depths[0, 0, 1] = 255
    quats = np.array([[0.1, 0.2, 0.3, 0.9],          # non-unit: scipy normalises
                      [0.0, 0.0, 0.0, 1.0],
                      [-0.3, 0.5, 0.1, 0.4]])
    trans = np.array([[0.5, -0.5, 1.0], [0.0, 0.0, 0.0], [1500.25, -3.5, 1720.125]])
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as td:
        os.chdir(td)
        try:
            for d in ("depth", "point", "point_world", "ply", "camera_pose"):
                os.mkdir(d)
            lines = ["id,tx,ty,tz,qx,qy,qz,qw,name,extra\n"]
            for k in range(n_frames):
                cv2.imwrite("depth/f%d.png" % k, depths[k])
                lines.append("%d,%r,%r,%r,%r,%r,%r,%r,f%d.png,0\n" % ((k,) + tuple(float(v) for v in trans[k]) +
                                                                  tuple(float(v) for v in quats[k]) + (k,)))
            pose_txt = "".join(lines)
            with open("camera_pose/image_colmap_simi_2.txt", "w") as f:
                f.write(pose_txt)
            c2w.main()
            cam_txt = [open("point/f%d.txt" % k).read() for k in range(n_frames)]
            world_txt_last = open("point_world/small_worldpoint_5_23_5.txt").read()
            ply_txt = open("ply/small_035_p8.ply").read()
        finally:
            os.chdir(cwd)
    # world points in float64, straight from the reference functions (no text in between)
    world = []
    rinvs = []
    for k in range(n_frames):
        r = c2w.scipy_transfer(quats[k])
        rinvs.append(np.asarray(r))
        pts = []
        for line in cam_txt[k].splitlines():
            p = c2w.str_tofloat(line.split(',')[0:3])
            pw = c2w.point_camera(p, r, trans[k])
            pts.append([pw[0, 0], pw[1, 0], pw[2, 0]])
        world.append(pts)
    world = np.array(world, dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "ref_c2w_small.npz"), depths=depths, quats=quats, trans=trans,
                        rinv=np.array(rinvs), world=world)
    with open(os.path.join(OUT, "ref_c2w_small_text.json"), "w") as f:
        json.dump({"pose_txt": pose_txt, "cam_txt": cam_txt, "world_txt_last": world_txt_last, "ply_txt": ply_txt}, f)
    meta["c2w_small"] = {"H": H, "W": W, "n_frames": n_frames, "ply_sha256": sha(ply_txt)}

    # ---- (2) pixel_to_camera.gentxtcord on a 480x640 frame (its loops are hard-coded to that size) ----
    d640 = (rng.integers(0, 256, size=(480, 640))).astype(np.uint8)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "p.txt")
        xyz = p2c.gentxtcord(path, d640)
        txt = open(path).read()
        ply_path = os.path.join(td, "p.ply")
        sub = [xyz[0][:50], xyz[1][:50], xyz[2][:50]]
        p2c.genply_RGB(sub, ply_path)
        ply50 = open(ply_path).read()
    sel = np.arange(0, 480 * 640, 997)
    np.savez_compressed(os.path.join(OUT, "ref_p2c_640.npz"), depth=d640, sel=sel,
                        X=np.array(xyz[0], dtype=np.float64)[sel], Y=np.array(xyz[1], dtype=np.float64)[sel],
                        Z=np.array(xyz[2], dtype=np.float64)[sel])
    meta["p2c_640"] = {"txt_sha256": sha(txt), "txt_first_lines": txt.splitlines()[:5], "ply50": ply50}

    # ---- (3) known-answer values quoted in SURVEY.md section 8c (regenerated, not transcribed) ----
    q = np.array([0.1, 0.2, 0.3, 0.9])
    r = c2w.scipy_transfer(q)
    pw = c2w.point_camera(np.zeros(3), r, np.array([0.5, -0.5, 1.0]))
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "z.txt")
        c2w.gentxtcord(path, np.zeros((1, 1), dtype=np.uint8))
        z0 = open(path).read()
    meta["kat"] = {"q": q.tolist(), "t": [0.5, -0.5, 1.0], "rinv": np.asarray(r).tolist(),
                   "world_of_origin": [str(pw[0, 0]), str(pw[1, 0]), str(pw[2, 0])], "gentxtcord_z0": z0}

    # ---- (4) imread behaviour the drop-in decode must mirror (a1): 16-bit PNG under IMREAD_GRAYSCALE ----
    with tempfile.TemporaryDirectory() as td:
        v16 = np.arange(65536, dtype=np.uint16).reshape(256, 256)
        cv2.imwrite(os.path.join(td, "v16.png"), v16)
        g = cv2.imread(os.path.join(td, "v16.png"), cv2.IMREAD_GRAYSCALE)
        meta["imread_gray_16bit_is_shift8"] = bool(np.array_equal(g, (v16 >> 8).astype(np.uint8)))
    # ---- (5) BASELINE config 1 at its real shape: camera_to_world.main() on ONE 1242x375 KITTI-shape depth PNG ----
    # (transfer/camera_to_world.py:178-180 -> get_file_name :138-174 -> gentxtcord :67-83, get_pointdata :86-105,
    # genply :112-134).  The frame is frame 0 of the synthetic C2 sequence (oracle/points_oracle.py, seed 20261018 + 1),
    # written as a 16-bit PNG; the reference reads it with IMREAD_GRAYSCALE, i.e. as uint8 = raw >> 8 (integer metres).
    # ~25 s of reference time.  Outputs are 19 / 26 / 12 MB of text: committed as sha256 + a strided sample.
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    from oracle import points_oracle as po
    W1, H1 = 1242, 375
    d16 = po.synth_depth_u16(W1, H1, po.KITTI_INTRINSICS, 20261018 + 1, "street")
    q1, t1 = po.synth_pose(0, 4500)
    with tempfile.TemporaryDirectory() as td:
        os.chdir(td)
        try:
            for d in ("depth", "point", "point_world", "ply", "camera_pose"):
                os.mkdir(d)
            cv2.imwrite("depth/000000.png", d16)
            pose_txt1 = "id,tx,ty,tz,qx,qy,qz,qw,name,extra\n" + "0,%r,%r,%r,%r,%r,%r,%r,000000.png,0\n" % (
                tuple(float(v) for v in t1) + tuple(float(v) for v in q1))
            with open("camera_pose/image_colmap_simi_2.txt", "w") as f:
                f.write(pose_txt1)
            c2w.main()
            cam1 = open("point/000000.txt").read()
            world1 = open("point_world/small_worldpoint_5_23_5.txt").read()
            ply1 = open("ply/small_035_p8.ply").read()
            seen8 = cv2.imread("depth/000000.png", cv2.IMREAD_GRAYSCALE)
        finally:
            os.chdir(cwd)
    assert np.array_equal(seen8, (d16 >> 8).astype(np.uint8))
    wl = world1.splitlines()
    sel1 = np.arange(0, W1 * H1, 97)
    world_sel = np.array([[float(v) for v in wl[i].split(',')] for i in sel1], dtype=np.float64)
    cl = cam1.splitlines()
    pl = ply1.split("end_header\n    ")[1].split("\n")
    np.savez_compressed(os.path.join(OUT, "ref_c1_kitti.npz"), depth8=seen8, q=q1, t=t1, rinv=np.asarray(c2w.scipy_transfer(q1)),
                        sel=sel1, world_sel=world_sel)
    meta["c1_kitti"] = {"W": W1, "H": H1, "seed": 20261018 + 1, "pose_txt": pose_txt1,
                        "depth16_sha256": sha(d16.tobytes()), "cam_txt_sha256": sha(cam1), "world_txt_sha256": sha(world1),
                        "ply_sha256": sha(ply1), "cam_txt_bytes": len(cam1), "world_txt_bytes": len(world1), "ply_bytes": len(ply1),
                        "cam_lines_sample": {str(i): cl[i] for i in (0, 1, 607, 232254, 465749)},
                        "world_lines_sample": {str(i): wl[i] for i in (0, 1, 607, 232254, 465749)},
                        "ply_rows_sample": {str(i): pl[i] for i in (0, 1, 607, 232254, 465749)}}
    with open(os.path.join(OUT, "ref_meta.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print("golden written to", os.path.abspath(OUT))
    print(json.dumps(meta["kat"], indent=1))
    print("imread 16-bit gray == >>8:", meta["imread_gray_16bit_is_shift8"])


if __name__ == "__main__":
    main()
