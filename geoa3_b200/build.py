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
SOURCES = ["nn_pair.cu", "knn.cu", "kappa_loss.cu", "pointnet2.cu"]
HEADERS = ["common.cuh", "csr.cuh", osp.join("..", "..", "include", "geoa3_b200.h")]
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
    if not force and not needs_build():
        return SO
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
        [osp.join(CSRC, s) for s in SOURCES] + ["-o", SO]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
