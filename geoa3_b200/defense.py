"""Point-removal defenses (same names, arguments and return convention as the reference's defense.py:18-50):
every function returns (filtered cloud [b,3,n'], number of removed points).

    random_drop_fn(pc, drop_num)                                         defense.py:18-23
    outlier_removal_fn(pc, defense_type, drop_num, alpha, outlier_knn)   defense.py:25-40
    point_removal_fn(pc, defense_type, drop_num, alpha, outlier_knn)     defense.py:42-50

The statistical outlier filters score every point by the mean distance to its `outlier_knn` nearest
neighbours.  The reference builds a dense [b,n,n] distance matrix and calls topk; here the neighbour search
runs on the exact top-K kernel (geoa3_knn) and only the n*k selected distances are formed.  A column
selection by a boolean keep-mask replaces the reference's sort-the-kept-indices step (identical result:
surviving points keep their original order).

Like the reference, 'outliers_variance' filters cloud 0 only and 'outliers_fixNum' expects b == 1 (the
evaluation script feeds one cloud at a time, defense.py:118-131)."""
import torch

from .loss_utils import _nbr_vectors, _self_nbr

_FILTERS = ("outliers_variance", "outliers_fixNum")


def _keep_columns(pc, keep):
    """pc [b,3,n], keep [n] bool -> the kept columns in their original order (a fresh contiguous tensor)."""
    return pc[:, :, keep.nonzero(as_tuple=True)[0]].contiguous()


def random_drop_fn(pc, drop_num):
    keep = torch.ones(pc.size(2), dtype=torch.bool, device=pc.device)
    keep[torch.randperm(pc.size(2))[:drop_num].to(pc.device)] = False
    return _keep_columns(pc, keep), drop_num


def knn_mean_distance(pc, outlier_knn):
    """Outlier score [b,n]: mean over the `outlier_knn` nearest neighbours of |p_j - p_i + 1e-10| — the reference
    adds the epsilon to every coordinate difference before squaring (:26), so it is kept inside the norm."""
    diff = _nbr_vectors(pc, _self_nbr(pc, outlier_knn)) + 1e-10
    return (diff * diff).sum(dim=1).sqrt().mean(dim=-1)


def outlier_removal_fn(pc, defense_type, drop_num, alpha, outlier_knn):
    if defense_type not in _FILTERS:
        raise AssertionError("Wrong defense type!")
    score = knn_mean_distance(pc, outlier_knn)
    n = pc.size(2)
    if defense_type == "outliers_variance":
        # keep what lies below mean + alpha*std of the cloud's own scores; cloud 0 only, as in the reference (:32-36)
        s0 = score[0]
        kept = _keep_columns(pc[:1], s0 < s0.mean() + alpha * s0.std())
        return kept, n - kept.size(2)
    # 'outliers_fixNum': the n - drop_num lowest scores survive (:37-40); topk picks the same set as the reference
    keep = torch.zeros(n, dtype=torch.bool, device=pc.device)
    keep[score.topk(n - drop_num, dim=1, largest=False, sorted=True)[1].reshape(-1)] = True
    kept = _keep_columns(pc, keep)
    return kept, n - kept.size(2)


def point_removal_fn(pc, defense_type, drop_num, alpha, outlier_knn):
    if defense_type == "rand_drop":
        return random_drop_fn(pc, drop_num)
    return outlier_removal_fn(pc, defense_type, drop_num, alpha, outlier_knn)
