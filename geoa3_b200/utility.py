"""Host-side helpers the GeoA3 loss path imports from the reference's Lib/utility.py."""
import torch


def _normalize(input, p=2, dim=1, eps=1e-12):
    """x / max(||x||_p, eps) along `dim` (reference: Lib/utility.py:30-31)."""
    return input / input.norm(p, dim, keepdim=True).clamp(min=eps).expand_as(input)


def _compare(output, target, gt, targeted):
    """Attack success predicate (reference: Lib/utility.py:151-156): targeted -> hit the target
    label, untargeted -> leave the ground-truth label."""
    if targeted:
        return output == target
    return output != gt


def farthest_points_sample(obj_points, num_points, start=None):
    """obj_points [b,3,n] -> the num_points farthest-point-sampled columns [b,3,num_points], first pick random
    (reference: Lib/utility.py:175-187, a Python loop of num_points-1 torch passes; here one kernel launch).
    Differentiable w.r.t. obj_points through the gather, like the reference.  `start` [b] fixes the first picks
    (tests); default = uniform random indices drawn on the device."""
    from . import ops

    assert obj_points.size(1) == 3
    b, _, n = obj_points.size()
    if start is None:
        start = torch.randint(n, (b,), device=obj_points.device, dtype=torch.int32)
    xyz = obj_points.detach().permute(0, 2, 1).contiguous().float()
    selected = ops.farthest_points_sample_idx(xyz, num_points, start.to(torch.int32).contiguous())
    return torch.gather(obj_points, 2, selected.long().unsqueeze(1).expand(b, 3, num_points))
