"""Host logic of the multi-GPU path on CPU: gloo, world_size 2 and 3 (one process per rank), see tests/workers/gloo_merge_worker.py."""
import importlib
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sharding = importlib.import_module("3d_reconstruction_system_b200.sharding")


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_frame_range_and_rounds():
    for n in (0, 1, 7, 4500, 10000):
        for w in (1, 2, 3, 8):
            edges = [sharding.frame_range(n, w, r) for r in range(w)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in edges) - min(b - a for a, b in edges) <= 1
            seen = []
            for base, parts in sharding.scan_rounds(n, w, 3):
                for r, a, c in parts:
                    seen.extend(range(a, a + c))
            assert seen == list(range(n))


def test_brick_owner_matches_kernel_hash(r3d):
    # same murmur3 finaliser as r3d_math.cuh hash64; spot values computed by hand from the definition
    keys = np.array([0, 1, 0x1fff | (0x1fff << 13) | (0x1fff << 26), 123456789], dtype=np.uint64)
    def h(x):
        m = (1 << 64) - 1
        x ^= x >> 33; x = (x * 0xff51afd7ed558ccd) & m; x ^= x >> 33; x = (x * 0xc4ceb9fe1a85ec53) & m; x ^= x >> 33
        return x
    for w in (1, 2, 8):
        assert list(sharding.brick_owner(keys, w)) == [(h(int(k)) >> 32) % w for k in keys]
    own = sharding.brick_owner(np.arange(100000, dtype=np.uint64), 8)
    assert np.bincount(own, minlength=8).min() > 11000      # balanced


@pytest.mark.parametrize("world,n_scans,per_rank,mode", [(2, 23, 3, "plain"), (3, 10, 2, "plain"), (2, 4, 4, "plain"), (2, 17, 2, "overlap"), (3, 19, 2, "regrow"), (2, 9, 3, "regrow")])
def test_scan_ordered_merge_gloo(world, n_scans, per_rank, mode):
    port = free_port()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "workers", "gloo_merge_worker.py"), str(n_scans), str(per_rank), mode],
                                      env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=240)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append(out)
    for r, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, "rank %d failed:\n%s" % (r, out)
        assert "rank %d ok" % r in out
