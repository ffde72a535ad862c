"""ctypes binding of libgeoa3_b200.so (the C ABI declared in include/geoa3_b200.h).

There is NO fallback: if the shared library is missing or a call fails, an exception is raised.
torch is used only for device memory and the current stream.
"""
import ctypes as C
import os
import os.path as osp

import torch

_HERE = osp.dirname(osp.abspath(__file__))
SO_PATH = os.environ.get("GEOA3_SO_PATH") or osp.join(_HERE, "libgeoa3_b200.so")  # override: A/B builds (tools only)

_vp, _i, _f, _sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t

# name -> (restype, argtypes): must list every symbol of include/geoa3_b200.h (checked by tests)
SIGNATURES = {
    "geoa3_version": (_i, []),
    "geoa3_error_string": (C.c_char_p, [_i]),
    "geoa3_nn_pair": (_i, [_vp, _vp, _i, _i, _i] + [_vp] * 11),
    "geoa3_knn": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp]),
    "geoa3_knn_set": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp]),
    "geoa3_cell_grid_max": (_i, [_i]),
    "geoa3_cell_blob_bytes": (_sz, [_i, _i]),
    "geoa3_cell_sort": (_i, [_vp, _i, _i, _i, _f, _i, _i, _i, _vp, _vp]),
    "geoa3_knn_cells": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _i, _vp, _vp, _vp]),
    "geoa3_nn_pair_cells": (_i, [_vp, _vp, _i, _i, _i, _i, _i] + [_vp] * 7),
    "geoa3_group_bbox_floats": (_sz, [_i]),
    "geoa3_group_bbox": (_i, [_vp, _i, _i, _vp, _vp]),
    "geoa3_arrange": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp]),
    "geoa3_kappa_loss_fwd": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "geoa3_loss_bwd_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "geoa3_loss_bwd": (_i, [_vp] * 13 + [_i, _i, _i, _i, _vp, _vp, _sz, _vp]),
    "geoa3_geo_fwd_bwd_supported": (_i, [_i, _i, _i]),
    "geoa3_geo_fwd_bwd": (_i, [_vp] * 7 + [_i, _vp, _vp, _f, _f, _f, _i, _i, _i] + [_vp] * 8),
    "geoa3_furthest_point_sampling": (_i, [_vp, _i, _i, _i, _vp, _vp]),
    "geoa3_farthest_points_sample": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp]),
    "geoa3_gather_points": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "geoa3_gather_points_grad": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _sz, _vp]),
    "geoa3_ball_query": (_i, [_vp, _vp, _i, _i, _i, _f, _i, _vp, _vp]),
    "geoa3_group_points": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "geoa3_group_points_grad_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "geoa3_group_points_grad": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _sz, _vp]),
    "geoa3_three_nn": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "geoa3_three_interpolate": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "geoa3_three_interpolate_grad": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _sz, _vp]),
}

_lib = None


def load():
    """Loads the library (once). Raises ImportError with build instructions when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not osp.exists(SO_PATH):
        raise ImportError(
            "geoa3_b200: %s not found — build it with `python -m geoa3_b200.build` (nvcc, sm_100a). "
            "There is no CPU/PyTorch fallback for the GeoA3 hot path." % SO_PATH)
    lib = C.CDLL(SO_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the ABI and this table diverge
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code):
    if code != 0:
        msg = load().geoa3_error_string(int(code))
        raise RuntimeError("libgeoa3_b200: error %d: %s" % (code, msg.decode() if msg else "?"))


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def stream(t):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def require_cuda_f32(t, name):
    """Same contract as the reference shims (_ext-src/include/utils.h:5-25), as a Python exception."""
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor (CPU not supported)" % name)
    if t.dtype != torch.float32:
        raise RuntimeError("%s must be a float tensor" % name)
    if not t.is_contiguous():
        raise RuntimeError("%s must be a contiguous tensor" % name)


def require_cuda_i32(t, name):
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor (CPU not supported)" % name)
    if t.dtype != torch.int32:
        raise RuntimeError("%s must be an int tensor" % name)
    if not t.is_contiguous():
        raise RuntimeError("%s must be a contiguous tensor" % name)
