"""Set-abstraction / feature-propagation modules on top of the B200 sampling+grouping ops.

Interface mirror of the reference's pointnet2_ops/pointnet2_modules.py (build_shared_mlp :10,
_PointnetSAModuleBase.forward :29-74, PointnetSAModuleMSG :77, PointnetSAModule :118,
PointnetFPModule :149): same constructor arguments, attribute names (`npoint`, `groupers`, `mlps`) and
state_dict layout, so reference checkpoints load.  The shared MLPs stay cuDNN (north star); only
FPS / gather / ball query / grouping run on libgeoa3_b200.so."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import pointnet2_utils as pu


def build_shared_mlp(mlp_spec, bn=True):
    mods = []
    for cin, cout in zip(mlp_spec[:-1], mlp_spec[1:]):
        mods.append(nn.Conv2d(cin, cout, kernel_size=1, bias=not bn))
        if bn:
            mods.append(nn.BatchNorm2d(cout))
        mods.append(nn.ReLU(True))
    return nn.Sequential(*mods)


class _PointnetSAModuleBase(nn.Module):
    def __init__(self):
        super().__init__()
        self.npoint, self.groupers, self.mlps = None, None, None

    def forward(self, xyz, features):
        """xyz (B,N,3), features (B,C,N)|None -> new_xyz (B,npoint,3)|None, new_features (B,sum mlp[-1],npoint)"""
        new_xyz = None
        if self.npoint is not None:
            centroid_idx = pu.furthest_point_sample(xyz, self.npoint)
            new_xyz = pu.gather_operation(xyz.transpose(1, 2).contiguous(), centroid_idx).transpose(1, 2).contiguous()
        pooled = []
        for grouper, mlp in zip(self.groupers, self.mlps):
            grouped = mlp(grouper(xyz, new_xyz, features))            # (B, mlp[-1], npoint, nsample)
            pooled.append(F.max_pool2d(grouped, kernel_size=[1, grouped.size(3)]).squeeze(-1))
        return new_xyz, torch.cat(pooled, dim=1)


class PointnetSAModuleMSG(_PointnetSAModuleBase):
    """Multi-scale grouping SA layer: one (radius, nsample, mlp) triple per scale."""

    def __init__(self, npoint, radii, nsamples, mlps, bn=True, use_xyz=True):
        super().__init__()
        assert len(radii) == len(nsamples) == len(mlps)
        self.npoint = npoint
        self.groupers, self.mlps = nn.ModuleList(), nn.ModuleList()
        for radius, nsample, spec in zip(radii, nsamples, mlps):
            self.groupers.append(pu.QueryAndGroup(radius, nsample, use_xyz=use_xyz) if npoint is not None
                                 else pu.GroupAll(use_xyz))
            if use_xyz:
                spec[0] += 3  # in place on the caller's list, like the reference (:107-108)
            self.mlps.append(build_shared_mlp(spec, bn))


class PointnetSAModule(PointnetSAModuleMSG):
    """Single-scale SA layer (npoint=None => group all)."""

    def __init__(self, mlp, npoint=None, radius=None, nsample=None, bn=True, use_xyz=True):
        super().__init__(mlps=[mlp], npoint=npoint, radii=[radius], nsamples=[nsample], bn=bn, use_xyz=use_xyz)


class PointnetFPModule(nn.Module):
    """Feature propagation by inverse-distance three-NN interpolation."""

    def __init__(self, mlp, bn=True):
        super().__init__()
        self.mlp = build_shared_mlp(mlp, bn=bn)

    def forward(self, unknown, known, unknow_feats, known_feats):
        if known is not None:
            dist, idx = pu.three_nn(unknown, known)
            recip = 1.0 / (dist + 1e-8)
            weight = recip / torch.sum(recip, dim=2, keepdim=True)
            interpolated = pu.three_interpolate(known_feats, idx, weight)
        else:
            interpolated = known_feats.expand(*(list(known_feats.size()[0:2]) + [unknown.size(1)]))
        feats = interpolated if unknow_feats is None else torch.cat([interpolated, unknow_feats], dim=1)
        return self.mlp(feats.unsqueeze(-1)).squeeze(-1)
