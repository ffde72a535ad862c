"""B200 drop-in for the reference's pointnet2_ops/pointnet2_utils.py.

Public surface (identical names, argument order, dtypes and autograd conventions; reference lines in brackets):

    furthest_point_sample(xyz (B,N,3) f32, npoint) -> (B,npoint) i32, non-differentiable        [:34-65]
    gather_operation(features (B,C,N), idx (B,npoint) i32) -> (B,C,npoint)                      [:68-101]
    three_nn(unknown (B,n,3), known (B,m,3)) -> (dist (B,n,3), idx (B,n,3) i32), non-diff.     [:104-136]
    three_interpolate(features (B,c,m), idx (B,n,3), weight (B,n,3)) -> (B,c,n)                [:139-191]
    grouping_operation(features (B,C,N), idx (B,npoint,nsample) i32) -> (B,C,npoint,nsample)   [:194-240]
    ball_query(radius, nsample, xyz (B,N,3), new_xyz (B,npoint,3)) -> (B,npoint,nsample) i32   [:243-276]
    QueryAndGroup(radius, nsample, use_xyz=True), GroupAll(use_xyz=True)                        [:279-379]

Note the Python order of ball_query's arguments differs from the native one (new_xyz, xyz, radius, nsample)
[:265].  Backward conventions kept: index-producing ops return `()`; GatherOperation returns (grad, None);
GroupingOperation / ThreeInterpolate return zero tensors for their index / weight inputs.  The native side
is libgeoa3_b200.so through `_ext`; gradients are deterministic gathers instead of float atomics.
There is no JIT build and no CPU path: CPU tensors raise RuntimeError.
"""
import torch
from torch import nn
from torch.autograd import Function

from . import _ext


class _IndexOp(Function):
    """Base of the ops whose outputs are indices: nothing to differentiate."""

    @staticmethod
    def backward(ctx, *unused_grads):
        return ()


class FurthestPointSampling(_IndexOp):
    @staticmethod
    def forward(ctx, xyz, npoint):
        picked = _ext.furthest_point_sampling(xyz, npoint)
        ctx.mark_non_differentiable(picked)
        return picked


class BallQuery(_IndexOp):
    @staticmethod
    def forward(ctx, radius, nsample, xyz, new_xyz):
        members = _ext.ball_query(new_xyz, xyz, radius, nsample)
        ctx.mark_non_differentiable(members)
        return members


class ThreeNN(_IndexOp):
    @staticmethod
    def forward(ctx, unknown, known):
        sq_dist, nearest = _ext.three_nn(unknown, known)
        dist = sq_dist.sqrt()
        ctx.mark_non_differentiable(dist, nearest)
        return dist, nearest


class GatherOperation(Function):
    @staticmethod
    def forward(ctx, features, idx):
        ctx.save_for_backward(idx)
        ctx.n_src = features.size(2)
        return _ext.gather_points(features, idx)

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        return _ext.gather_points_grad(grad_out.contiguous(), idx, ctx.n_src), None


class GroupingOperation(Function):
    @staticmethod
    def forward(ctx, features, idx):
        ctx.save_for_backward(idx)
        ctx.n_src = features.size(2)
        return _ext.group_points(features, idx)

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        return _ext.group_points_grad(grad_out.contiguous(), idx, ctx.n_src), torch.zeros_like(idx)


class ThreeInterpolate(Function):
    @staticmethod
    def forward(ctx, features, idx, weight):
        ctx.save_for_backward(idx, weight)
        ctx.n_src = features.size(2)
        return _ext.three_interpolate(features, idx, weight)

    @staticmethod
    def backward(ctx, grad_out):
        idx, weight = ctx.saved_tensors
        grad = _ext.three_interpolate_grad(grad_out.contiguous(), idx, weight, ctx.n_src)
        return grad, torch.zeros_like(idx), torch.zeros_like(weight)


furthest_point_sample = FurthestPointSampling.apply
gather_operation = GatherOperation.apply
three_nn = ThreeNN.apply
three_interpolate = ThreeInterpolate.apply
grouping_operation = GroupingOperation.apply
ball_query = BallQuery.apply


def _stack_xyz_and_features(grouped_xyz, grouped_features, use_xyz):
    if grouped_features is None:
        return grouped_xyz
    return torch.cat([grouped_xyz, grouped_features], dim=1) if use_xyz else grouped_features


class QueryAndGroup(nn.Module):
    """Ball query around `new_xyz`, then gather coordinates (relative to the centroid) and features:
    (B,N,3), (B,npoint,3), (B,C,N)|None -> (B, 3+C, npoint, nsample)."""

    def __init__(self, radius, nsample, use_xyz=True):
        super().__init__()
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz

    def forward(self, xyz, new_xyz, features=None):
        members = ball_query(self.radius, self.nsample, xyz, new_xyz)
        rel = grouping_operation(xyz.transpose(1, 2).contiguous(), members)
        rel -= new_xyz.transpose(1, 2).unsqueeze(-1)  # in place on the op's output, like the reference [:317]
        if features is None:
            assert self.use_xyz, "Cannot have not features and not use xyz as a feature!"
            return rel
        return _stack_xyz_and_features(rel, grouping_operation(features, members), self.use_xyz)


class GroupAll(nn.Module):
    """One group holding every point: (B,N,3), _, (B,C,N)|None -> (B, 3+C, 1, N).  No native op involved."""

    def __init__(self, use_xyz=True):
        super().__init__()
        self.use_xyz = use_xyz

    def forward(self, xyz, new_xyz, features=None):
        everything = xyz.transpose(1, 2).unsqueeze(2)
        return _stack_xyz_and_features(everything, None if features is None else features.unsqueeze(2), self.use_xyz)
