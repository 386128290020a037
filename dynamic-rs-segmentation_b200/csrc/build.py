"""Build libdrs.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python dynamic-rs-segmentation_b200/csrc/build.py [--force] [--verbose]

The shared object is written next to the package (``dynamic-rs-segmentation_b200/libdrs.so``) so that it
travels with the repo snapshot to the GPU box; it is git-ignored.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
OUT = os.path.join(PKG, "libdrs.so")
STAMP = os.path.join(PKG, ".libdrs.stamp")
SOURCES = ["drs_api.cu"]
DEPS = ["drs_api.cu", "drs_train.cuh", "drs_scene_api.cuh", "drs_common.cuh", "ptx_sm100.cuh", "conv_tc.cuh",
        "conv_simt.cuh", "ops.cuh", "scene.cuh", "wgrad_tc.cuh", "conv1_tc.cuh", "../../include/drs.h"]
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-shared", "-Xcompiler",
         "-fPIC", "-Xcompiler", "-fvisibility=default", "--expt-relaxed-constexpr", "-cudart", "static"]


def _digest():
    h = hashlib.sha256()
    for d in DEPS:
        p = os.path.join(HERE, d)
        if os.path.exists(p):
            with open(p, "rb") as f:
                h.update(f.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    dig = _digest()
    if not force and os.path.exists(OUT) and os.path.exists(STAMP) and open(STAMP).read().strip() == dig:
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + [os.path.join(HERE, s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libdrs.so")
    with open(STAMP, "w") as f:
        f.write(dig)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
