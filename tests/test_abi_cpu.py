"""CPU suite: the C-ABI library builds for sm_100a, loads, and exports exactly the symbols that
include/geoa3_b200.h declares (no compute calls without a GPU)."""
import ctypes
import os.path as osp
import re

import pytest

ROOT = osp.dirname(osp.dirname(osp.abspath(__file__)))


def _declared():
    src = open(osp.join(ROOT, "include", "geoa3_b200.h")).read()
    return sorted(set(re.findall(r"GEOA3_API[^;(]*?\b(geoa3_\w+)\s*\(", src)))


def test_library_builds_and_exports_header_symbols():
    from geoa3_b200 import _lib, build

    so = build.build()
    assert osp.exists(so)
    names = _declared()
    assert len(names) >= 16
    lib = ctypes.CDLL(so)
    for n in names:
        assert hasattr(lib, n), "missing export " + n
    assert sorted(_lib.SIGNATURES) == names, "ctypes table and header diverge"
    loaded = _lib.load()
    assert loaded.geoa3_version() >= 1000
    assert b"workspace" in loaded.geoa3_error_string(-3)


def test_argument_errors_do_not_need_a_gpu():
    from geoa3_b200 import _lib

    lib = _lib.load()
    assert lib.geoa3_nn_pair(None, None, 1, 8, 8, *([None] * 11)) == -1
    assert lib.geoa3_knn(None, None, 0, 0, 0, 1, 0, None, None, None, None, None, 0, None, None, None) == -1
    assert lib.geoa3_group_points_grad_workspace_bytes(2, 10, 4, 3) == 16 + 2 * (10 + 1 + 12) * 4


def test_cpu_tensors_fail_loudly():
    import torch

    from geoa3_b200 import loss_utils
    from geoa3_b200.pointnet2_ops import pointnet2_utils as pu

    a = torch.zeros(1, 3, 16)
    with pytest.raises(RuntimeError):
        loss_utils.chamfer_loss(a, a)
    with pytest.raises(RuntimeError):
        pu.furthest_point_sample(torch.zeros(1, 16, 3), 4)


def test_no_product_import_of_oracle():
    import glob

    for f in glob.glob(osp.join(ROOT, "geoa3_b200", "**", "*.py"), recursive=True):
        src = open(f).read()
        assert "import oracle" not in src and "from oracle" not in src, f
