"""Point-removal defenses of the reference's defense.py (:18-50) — the statistical outlier filters are kNN
statistics, so their O(n^2) part runs on the exact top-K kernel (geoa3_knn) instead of a dense [b,n,n]
matrix + topk.  Same names, arguments and return convention: (filtered cloud, number of removed points).

    random_drop_fn(pc, drop_num)                                         defense.py:18-23
    outlier_removal_fn(pc, defense_type, drop_num, alpha, outlier_knn)   defense.py:25-40
    point_removal_fn(pc, defense_type, drop_num, alpha, outlier_knn)     defense.py:42-50

Clouds are float32 CUDA tensors [b,3,n]; like the reference, 'outliers_variance' filters cloud 0 only and
'outliers_fixNum' expects b == 1 (the evaluation script feeds one cloud at a time, defense.py:118-131)."""
import torch

from .loss_utils import _nbr_vectors, _self_nbr


def random_drop_fn(pc, drop_num):
    n = pc.size(2)
    idx = torch.randperm(n)[drop_num:].long().to(pc.device)
    idx = torch.sort(idx, dim=0, descending=False)[0]
    return pc.clone()[:, :, idx].contiguous(), drop_num


def knn_mean_distance(pc, outlier_knn):
    """Mean distance to the `outlier_knn` nearest neighbours, measured as the reference does (:26-27):
    |p_j - p_i + 1e-10| with the epsilon added to every coordinate difference."""
    v = _nbr_vectors(pc, _self_nbr(pc, outlier_knn))
    return (v + 1e-10).pow(2).sum(dim=1).sqrt().mean(dim=-1)


def outlier_removal_fn(pc, defense_type, drop_num, alpha, outlier_knn):
    dis = knn_mean_distance(pc, outlier_knn)
    n = pc.size(2)
    if defense_type == 'outliers_variance':
        keep_mask = dis < (dis.mean(-1) + alpha * dis.std(-1)).unsqueeze(-1)
        output_pc = torch.masked_select(pc[0], keep_mask[0].unsqueeze(0).expand_as(pc[0])).view(1, 3, -1)
        return output_pc, pc.size(2) - output_pc.size(2)
    elif defense_type == 'outliers_fixNum':
        idx = dis.topk(n - drop_num, dim=1, largest=False, sorted=True)[1].view(-1)
        idx = torch.sort(idx, dim=0, descending=False)[0]
        return pc.clone()[:, :, idx].contiguous(), n - idx.size(0)
    assert False, 'Wrong defense type!'


def point_removal_fn(pc, defense_type, drop_num, alpha, outlier_knn):
    if defense_type == 'rand_drop':
        output_pc, num = random_drop_fn(pc, drop_num)
    elif defense_type == 'outliers_variance' or defense_type == 'outliers_fixNum':
        output_pc, num = outlier_removal_fn(pc, defense_type, drop_num, alpha, outlier_knn)
    else:
        assert False, 'Wrong defense type!'
    return output_pc, num
