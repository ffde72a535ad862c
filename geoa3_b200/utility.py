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


# ---------------------------------------------------------------------------- neighbourhood PCA estimators
def _centred_neighbours(pc, k):
    """pc [b,3,n] -> coordinates of the k nearest neighbours of every point (own entry dropped), centred on their
    mean: [b,n,3,k].  The O(n^2) search runs on the exact top-K kernel (the reference uses pytorch3d knn_points +
    knn_gather, Lib/utility.py:45-46,125-126)."""
    from . import ops

    b, _, n = pc.shape
    c = pc.detach().float().contiguous()
    nbr = ops.knn(c, c, int(k) + 1, drop=1)[0].long()
    pts = torch.gather(c, 2, nbr.reshape(b, 1, n * k).expand(b, 3, n * k)).view(b, 3, n, k).permute(0, 2, 1, 3)
    return pts - pts.mean(dim=3, keepdim=True)


def _cov_eigh(pc, k):
    """Per-point neighbourhood covariance (factor 1/(k-1)) and its eigen-decomposition, batched over all b*n points
    (the reference loops over the batch and calls the since-removed torch.symeig; eigenvalues ascending)."""
    cs = _centred_neighbours(pc, k)
    cov = (1.0 / (k - 1)) * torch.matmul(cs, cs.transpose(2, 3))
    w, v = torch.linalg.eigh(cov)
    return cs, w, v


def estimate_normal(pc, k):
    """Unit normals [b,3,n] = eigenvector of the smallest neighbourhood-covariance eigenvalue, with the reference's
    sign rule (Lib/utility.py:40-70; note that rule reads the sum of ALREADY CENTRED neighbours, i.e. rounding
    noise — the orientation is as arbitrary here as it is there)."""
    with torch.no_grad():
        cs, w, v = _cov_eigh(pc, k)
        nrm = v[..., :, 0]
        sign = -torch.sign((nrm * cs.sum(dim=3)).sum(-1, keepdim=True))
        return (sign * nrm).permute(0, 2, 1).contiguous().float()


def estimate_normal_via_ori_normal(pc_adv, pc_ori, normal_ori, k):
    """Normals for adversarial points borrowed from the original cloud (Lib/utility.py:92-110): the normal of the
    nearest original point where the point has not moved (squared distance < 1e-6), otherwise the normalised mean
    of the k nearest original normals."""
    from . import ops

    b, _, n = pc_adv.size()
    idx, dist = ops.knn(pc_adv.detach().float().contiguous(), pc_ori.detach().float().contiguous(), int(k),
                        return_dist=True)
    idx = idx.long()
    normal_pts = torch.gather(normal_ori, 2, idx.reshape(b, 1, n * k).expand(b, 3, n * k)).view(b, 3, n, k)
    avg = normal_pts.mean(dim=-1)
    avg = avg / (avg.norm(dim=1, keepdim=True) + 1e-12)
    first = normal_pts[:, :, :, 0]
    condition = (dist[:, :, 0] < 1e-6).unsqueeze(1).expand_as(first)
    return torch.where(condition, first, avg)


def jitter_input(data, sigma=0.01, clip=0.05):
    assert data.size(1) == 3 and clip > 0
    return torch.clamp(sigma * torch.randn_like(data), -1 * clip, clip)


def get_perpendicular_jitter(vector, sigma=0.01, clip=0.05):
    a1, a2 = sigma * torch.randn_like(vector), sigma * torch.randn_like(vector)
    return (torch.clamp(torch.cross(vector, a1, dim=1), -1 * clip, clip)
            + torch.clamp(torch.cross(vector, a2, dim=1), -1 * clip, clip))


def estimate_perpendicular(pc, k, sigma=0.01, clip=0.05):
    """Random jitter inside each point's tangent plane (Lib/utility.py:119-149): the two eigenvectors with the
    larger eigenvalues, each scaled by an independent N(0, sigma) draw per point and clipped."""
    with torch.no_grad():
        b, _, n = pc.size()
        _, w, v = _cov_eigh(pc, k)
        t1 = v[..., :, 2].permute(0, 2, 1)
        t2 = v[..., :, 1].permute(0, 2, 1)
        a1 = sigma * torch.randn(b, 1, n, device=pc.device)
        a2 = sigma * torch.randn(b, 1, n, device=pc.device)
        return torch.clamp(t1 * a1, -1 * clip, clip) + torch.clamp(t2 * a2, -1 * clip, clip)
