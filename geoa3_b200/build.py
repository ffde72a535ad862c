"""Builds geoa3_b200/libgeoa3_b200.so in-tree with nvcc for sm_100a (B200) only.

    python -m geoa3_b200.build [--force] [--verbose]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import os
import os.path as osp
import subprocess
import sys

HERE = osp.dirname(osp.abspath(__file__))
CSRC = osp.join(HERE, "csrc")
SO = osp.join(HERE, "libgeoa3_b200.so")
SOURCES = ["nn_pair.cu", "knn.cu", "knn_select.cu", "knn_cells.cu", "nn_cells.cu", "kappa_loss.cu", "pointnet2.cu"]
HEADERS = ["common.cuh", "csr.cuh", "cells.cuh", osp.join("..", "..", "include", "geoa3_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "--expt-relaxed-constexpr", "--extended-lambda",
    "-shared",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (osp.isabs(cand) and osp.exists(cand) or not osp.isabs(cand)):
            return cand
    return "nvcc"


def needs_build():
    if not osp.exists(SO):
        return True
    t = osp.getmtime(SO)
    deps = [osp.join(CSRC, s) for s in SOURCES + HEADERS] + [osp.abspath(__file__)]
    return any(osp.exists(d) and osp.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Every source is compiled to an object in parallel (one nvcc per file), then linked into the .so."""
    if not force and not needs_build():
        return SO
    objdir = osp.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    flags = [f for f in NVCC_FLAGS if f != "-shared"] + (["-Xptxas", "-v"] if verbose else [])
    flags += os.environ.get("GEOA3_NVCC_DEFS", "").split()  # experiment knobs (-DNAME=value), never set in production
    hdr_t = max(osp.getmtime(osp.join(CSRC, h)) for h in HEADERS if osp.exists(osp.join(CSRC, h)))
    jobs, objs = [], []
    for src in SOURCES:
        obj = osp.join(objdir, src.replace(".cu", ".o"))
        objs.append(obj)
        srcp = osp.join(CSRC, src)
        if not force and osp.exists(obj) and osp.getmtime(obj) > max(osp.getmtime(srcp), hdr_t, osp.getmtime(osp.abspath(__file__))):
            continue
        cmd = [_nvcc()] + flags + ["-c", srcp, "-o", obj]
        if verbose:
            print(" ".join(cmd))
        jobs.append((src, subprocess.Popen(cmd)))
    failed = [src for src, p in jobs if p.wait() != 0]
    if failed:
        raise RuntimeError("nvcc failed for " + ", ".join(failed))
    subprocess.check_call([_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", SO] + objs)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
