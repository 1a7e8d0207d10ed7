"""Small end-to-end case for compute-sanitizer (memcheck / racecheck): K1, K2, K3 (dense and hash), K4, K5, K6 once each."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
r3d = importlib.import_module("3d_reconstruction_system_b200")
octomap = importlib.import_module("3d_reconstruction_system_b200.octomap")

ctx = r3d.default_context(0)
rng = np.random.default_rng(0)
depth = rng.integers(0, 4000, size=(3, 64, 300)).astype(np.uint16)
q = np.array([[0.0, 0.0, 0.0, 1.0], [0.1, 0.0, 0.0, 0.9], [0.0, 0.2, 0.0, 0.9]])
t = np.array([[0.0, 0.0, 0.0], [0.5, 0.0, 0.1], [1.0, 0.1, 0.0]])
rt = ctx.pose_to_rt(q, t)
world, _ = ctx.backproject(depth, (300.0, 300.0, 150.0, 32.0), rt=rt, depth_scale=1 / 256.0)
comp, cnt = ctx.backproject(depth, (300.0, 300.0, 150.0, 32.0), rt=rt, depth_scale=1 / 256.0, compact=True)
tree = octomap.OcTree(0.1)
tree.updateNodes(world[:5000], True)
dev = ctx.to_device(world)
origins = np.zeros((3, 3))
tree.insertPointClouds(dev, origins, maxrange=12.0)                    # dense, pipelined
tree.insertPointCloud(world[:4000], origins[0], maxrange=-1.0)         # hash table
rec = tree.computeScanDelta(world[:2000], origins[0], maxrange=5.0)
tree.applyDelta(rec)
bt = tree.writeBinary()
other = octomap.OcTree(0.1)
other.readBinary(bt)
assert other.writeBinary() == bt
ply = ctx.ply_rows(world[:3000].astype(np.float64))
txt = ctx.txt_rows(world[:3000].astype(np.float64))
print("ok", world.shape, int(cnt.sum()), tree.numVoxels(), len(bt), len(ply), len(txt))
