"""TEST INFRASTRUCTURE ONLY — builds the *unmodified* reference pointnet2_ops
CUDA/C++ sources (where they lie under /root/reference) into oracle/_ref/ for
sm_100, so the GPU parity tests and the "reference kernel recompiled for
sm_100" timing bar can run the reference's own kernels on the B200 box.

Nothing is copied into the repo: the sources are compiled in place
(/root/reference/Model/pointnet2_ops_lib/pointnet2_ops/_ext-src) and only the
resulting pointnet2_ref_ext.so lands in oracle/_ref/ (git-ignored, shipped to the
GPU box by gpurun).  The reference's own JIT path is not used because it forces
an arch list starting at sm_37 (pointnet2_utils.py:23).

Usage:  python oracle/build_ref.py         (needs /root/reference; ~1 min)
"""
import glob
import os
import os.path as osp
import sys

REF_SRC = "/root/reference/Model/pointnet2_ops_lib/pointnet2_ops/_ext-src"
OUT_DIR = osp.join(osp.dirname(osp.abspath(__file__)), "_ref")
NAME = "pointnet2_ref_ext"


def built_path():
    return osp.join(OUT_DIR, NAME + ".so")


def build(verbose=False):
    """Compile the reference extension if its sources are present. Returns the
    .so path or None when /root/reference is absent (GPU box: prebuilt only)."""
    if not osp.isdir(REF_SRC):
        return built_path() if osp.exists(built_path()) else None
    if osp.exists(built_path()):
        return built_path()
    os.makedirs(OUT_DIR, exist_ok=True)
    os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0"
    os.environ.setdefault("MAX_JOBS", "8")
    from torch.utils.cpp_extension import load

    srcs = sorted(glob.glob(osp.join(REF_SRC, "src", "*.cpp")) + glob.glob(osp.join(REF_SRC, "src", "*.cu")))
    load(
        NAME,
        sources=srcs,
        extra_include_paths=[osp.join(REF_SRC, "include")],
        extra_cflags=["-O3"],
        extra_cuda_cflags=["-O3"],
        build_directory=OUT_DIR,
        with_cuda=True,
        is_python_module=False,
        verbose=verbose,
    )
    return built_path()


def load_ref():
    """Import the prebuilt reference extension (python module with the 9 ops of
    bindings.cpp:6-19). Returns None if it was never built."""
    p = built_path()
    if not osp.exists(p):
        return None
    import importlib.util
    import torch  # noqa: F401  (libtorch symbols must be loaded first)

    spec = importlib.util.spec_from_file_location(NAME, p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
