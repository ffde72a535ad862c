import glob
import os.path as osp

import numpy as np

GOLDEN_DIR = osp.join(osp.dirname(osp.abspath(__file__)), "golden")


def golden_files(prefix="loss_"):
    return sorted(glob.glob(osp.join(GOLDEN_DIR, prefix + "*.npz")))


def rel_err(a, b):
    """norm-wise relative error: max|a-b| / max|b|"""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def near_tie_rows(dist_sorted, ulps=2):
    """rows whose (K-1)-th/K-th boundary or internal order is decided by <= `ulps` ulp (index parity whitelist)"""
    d = np.asarray(dist_sorted, np.float32)
    gap = np.diff(d, axis=-1)
    tol = ulps * np.spacing(np.abs(d[..., 1:]))
    return (gap <= tol).any(-1)


def elem_err(a, b, rtol=1e-5, floor=1e-2):
    """Worst ELEMENT-WISE violation ratio of  |a-b| <= rtol*|b| + rtol*floor*max|b|  (<= 1 passes): every entry is held
    to `rtol` relative to ITSELF, down to entries `floor` times the largest one; below that the bound is absolute
    (rtol*floor*max|b|), because an fp32 sum of O(max|b|) terms cannot be resolved further."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    bound = rtol * np.abs(b) + rtol * floor * np.abs(b).max()
    return float((np.abs(a - b) / np.maximum(bound, 1e-300)).max())
