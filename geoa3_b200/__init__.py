"""geoa3_b200 — B200 (sm_100a) implementation of GeoA3's geometry-aware loss path and pointnet2_ops."""
import sys


def install_as_pointnet2_ops():
    """Registers this package's drop-in under the reference's import name, so that the reference's own callers
    (`from pointnet2_ops.pointnet2_modules import PointnetSAModule`, Model/PointNetPP_ssg.py:5, PointNetPP_msg.py:4;
    `import pointnet2_ops.pointnet2_utils`) resolve to the B200 ops without editing them.  Returns the package."""
    from . import pointnet2_ops as pkg
    from .pointnet2_ops import _ext, pointnet2_modules, pointnet2_utils

    sys.modules["pointnet2_ops"] = pkg
    sys.modules["pointnet2_ops._ext"] = _ext
    sys.modules["pointnet2_ops.pointnet2_utils"] = pointnet2_utils
    sys.modules["pointnet2_ops.pointnet2_modules"] = pointnet2_modules
    return pkg
