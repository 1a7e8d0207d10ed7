#!/usr/bin/env python
"""Pin the occupancy oracle against the REAL OctoMap library -- TEST INFRASTRUCTURE ONLY.

The occupancy arithmetic of the reference (OcTree / updateNode / insertPointCloud / writeBinary, SURVEY.md section 8
a9-a13) lives in the third-party `octomap` Python extension (wkentaro/octomap-python or neka-nat/python-octomap over
OctoMap 1.8 / 1.9), which is neither vendored in the reference nor installable in the authoring image.  This script is
the one command a maintainer runs on any machine where `import octomap` works:

    pip install octomap-python          # (or python-octomap)
    python oracle/pin_against_octomap.py            # -> tests/golden/octomap_pin.json
    python -m pytest tests/test_oracle_octomap.py -k pinned

It replays oracle/octomap_corpus.py (key boundaries, the log-odds ladder, the DDA tie / border cases, the .bt known
answers, random clouds, KITTI-shape scans of the benchmark) through the real module and records per case: size(), the
float32 bit patterns of getLogOdds() at the probes, and sha256 + length of writeBinary().  With that file present
`tests/test_oracle_octomap.py::test_pinned_against_real_octomap` compares the C restatement with it, and the GPU tests
(which compare the kernels with the restatement bit for bit) inherit the pin.

Without the module:  --self-check replays the corpus through the C restatement instead and writes to a scratch path,
which only proves that this script and the test agree on the corpus (it is what CI runs here).
"""
import argparse
import hashlib
import json
import os
import struct
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(ROOT, "tests", "golden", "octomap_pin.json")


def f32bits(v):
    return None if v is None else struct.unpack("<I", struct.pack("<f", float(v)))[0]


class RealTree:
    """The calls the reference scripts make, on the real extension (both bindings spell them the same way)."""

    def __init__(self, mod, res):
        self.t = mod.OcTree(float(res))

    def update(self, p, v):
        self.t.updateNode(np.asarray(p, dtype=np.float64), v)

    def insert(self, pts, origin, maxrange, discretize):
        self.t.insertPointCloud(np.ascontiguousarray(pts, dtype=np.float64), np.ascontiguousarray(origin, dtype=np.float64),
                                maxrange=float(maxrange), lazy_eval=False, discretize=bool(discretize))

    def logodds(self, p):
        try:
            node = self.t.search(np.asarray(p, dtype=np.float64))
            return float(node.getLogOdds())
        except Exception:          # both bindings raise (NullPointerException) for an unknown voxel
            return None

    def size(self):
        return int(self.t.size())

    def bt(self):
        self.t.updateInnerOccupancy()
        try:
            data = self.t.writeBinary()                      # octomap-python: no argument -> bytes
            if isinstance(data, (bytes, bytearray)):
                return bytes(data)
        except TypeError:
            pass
        with tempfile.TemporaryDirectory() as td:            # python-octomap: file only
            path = os.path.join(td, "t.bt")
            if not self.t.writeBinary(bytes(path, encoding="utf-8")):
                raise RuntimeError("writeBinary failed")
            return open(path, "rb").read()


class OracleTree:
    def __init__(self, mod, res):
        self.t = mod.OcTree(float(res))

    def update(self, p, v):
        self.t.updateNode(np.asarray(p, dtype=np.float64), v)

    def insert(self, pts, origin, maxrange, discretize):
        self.t.insertPointCloud(pts, origin, maxrange=float(maxrange), discretize=bool(discretize))

    def logodds(self, p):
        k = self.t.coordToKey(np.asarray(p, dtype=np.float64).astype(np.float32).astype(np.float64))
        return None if k is None else self.t.search(k)

    def size(self):
        return self.t.size()

    def bt(self):
        self.t.updateInnerOccupancy()
        return self.t.write_binary_bytes()


def replay(case, make_tree):
    """Run one corpus case on a fresh tree -> the record stored in / compared with the golden file."""
    tree = make_tree(case["res"])
    for op in case["ops"]:
        if op[0] == "update":
            for p in op[1]:
                tree.update(p, op[2])
        else:
            tree.insert(op[1], op[2], op[3], op[4])
    lo = [f32bits(tree.logodds(p)) for p in case["probes"]]
    size = tree.size()
    bt = tree.bt()
    return {"size": size, "logodds_bits": lo, "bt_len": len(bt), "bt_sha256": hashlib.sha256(bt).hexdigest(),
            "bt_head": bt[:200].decode("latin-1") if len(bt) < 400 else None}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=GOLDEN)
    ap.add_argument("--self-check", action="store_true", help="replay through the C restatement (plumbing check only; never writes the golden path)")
    ap.add_argument("--light", action="store_true", help="skip the KITTI-shape scans (a minute of CPU on the real library)")
    args = ap.parse_args()
    sys.path.insert(0, ROOT)
    from oracle import octomap_corpus
    if args.self_check:
        from oracle import octomap_oracle as mod
        make = lambda res: OracleTree(mod, res)
        info = {"library": "oracle/octomap_oracle.c (SELF-CHECK, not a pin)"}
        if os.path.abspath(args.out) == os.path.abspath(GOLDEN):
            args.out = os.path.join(tempfile.gettempdir(), "octomap_pin_selfcheck.json")
    else:
        # the repo root holds an `octomap/` directory of drop-in scripts: make sure the extension module is what gets imported
        sys.path = [p for p in sys.path if os.path.abspath(p or ".") != ROOT]
        try:
            import octomap as mod
        except ImportError as exc:
            raise SystemExit("pin_against_octomap: `import octomap` failed (%s).  Install octomap-python (or python-octomap) and run again; "
                             "the occupancy oracle stays 'parity unpinned' until then." % exc)
        if not hasattr(mod, "OcTree"):
            raise SystemExit("pin_against_octomap: the imported `octomap` (%s) is not the OctoMap extension" % getattr(mod, "__path__", mod))
        make = lambda res: RealTree(mod, res)
        info = {"library": "octomap extension", "module_file": getattr(mod, "__file__", None), "version": getattr(mod, "__version__", None)}
    out = {"info": info, "cases": {}}
    for case in octomap_corpus.cases(heavy=not args.light):
        out["cases"][case["name"]] = replay(case, make)
        print("%-36s size %8d  bt %9d B  %s" % (case["name"], out["cases"][case["name"]]["size"], out["cases"][case["name"]]["bt_len"],
                                                 out["cases"][case["name"]]["bt_sha256"][:16]))
    with open(args.out, "w") as f:
        json.dump(out, f, indent=1)
    print("written", args.out)


if __name__ == "__main__":
    main()
