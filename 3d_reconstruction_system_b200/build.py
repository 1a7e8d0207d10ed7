"""In-tree build of libr3d_b200.so (hand-written CUDA for sm_100a + the C ABI of include/r3d.h).

    python -m 3d_reconstruction_system_b200.build   (or __graft_entry__.build())

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels with the tree to the GPU box.
"""
import concurrent.futures
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# R3D_BUILD_TAG=<tag> (kernel experiments, together with R3D_NVCC_EXTRA) builds libr3d_b200_<tag>.so beside the product
# library from its own object directory; R3D_LIB_PATH selects it at load time (_lib.py).  The default build is untagged.
TAG = os.environ.get("R3D_BUILD_TAG", "")
OBJ = os.path.join(CSRC, "_obj" + ("_" + TAG if TAG else ""))
LIB = os.path.join(HERE, "libr3d_b200%s.so" % ("_" + TAG if TAG else ""))

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",                      # no FMA contraction: rounding order is part of the parity contract
    "-Xcompiler", "-fPIC,-ffp-contract=off,-O2",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built (there is no CPU fallback)")


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_mtime():
    m = 0.0
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for f in os.listdir(root):
            if f.endswith((".cu", ".cuh", ".h")):
                m = max(m, os.path.getmtime(os.path.join(root, f)))
    return max(m, os.path.getmtime(__file__))


def _compile(src, log):
    obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
    extra = os.environ.get("R3D_NVCC_EXTRA", "").split()      # e.g. -DK3_VARIANT=1 for kernel experiments
    cmd = [_nvcc()] + NVCC_FLAGS + extra + ["-c", src, "-o", obj]
    p = subprocess.run(cmd, capture_output=True, text=True)
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + p.stdout + p.stderr)
    if p.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s" % (src, (p.stdout + p.stderr)[-6000:]))
    return obj


def build(force=False, verbose=False):
    """Compile every csrc/*.cu for sm_100a and link libr3d_b200.so.  Returns the library path."""
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _deps_mtime():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    srcs = sources()
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, os.path.join(OBJ, os.path.basename(s) + ".log")), srcs))
    cmd = [_nvcc(), "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart_static", "-lpthread", "-ldl", "-lrt", "-lz"]
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError("link failed:\n" + p.stdout + p.stderr)
    if verbose:
        for s in srcs:
            print(open(os.path.join(OBJ, os.path.basename(s) + ".log")).read())
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
