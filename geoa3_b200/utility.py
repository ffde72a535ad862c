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
