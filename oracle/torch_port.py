"""TEST / BASELINE INFRASTRUCTURE ONLY — dense PyTorch (CPU) port of the reference loss path, used as the
`cpu_baseline` / `--impl reference` arm of bench.py (kind "port": the reference's own Lib/loss_utils.py is
Python and does not exist on the GPU box, so it cannot be executed there).

It restates Lib/loss_utils.py:28-97 with the third-party kNN (pytorch3d, un-vendored) replaced by the dense
formulation the reference itself documents in comments: squared-L2 matrix + topk(K, largest=False,
sorted=True) (loss_utils.py:30-31,46-47,54-56,67-69,74-76), i.e. exactly the configuration BASELINE.md §3
prescribes for the CPU baseline.  Validated against the reference run verbatim in tests/test_port_cpu.py.
The attack step mirrors Attacker/geoA3_attack.py:100-180,319-329 (CE untargeted, CD + 0.1 HD + curvature, Adam).
"""
import time

import torch


def _knn_dense(p1, p2, K):
    """p1 [b,3,n], p2 [b,3,m] -> dists [b,n,K], idx [b,n,K] (ascending)"""
    d = ((p1.unsqueeze(3) - p2.unsqueeze(2)) ** 2).sum(1)
    return torch.topk(d, K, dim=2, largest=False, sorted=True)


def _normalize(x, eps=1e-12):  # Lib/utility.py:30-31
    return x / x.norm(2, 1, keepdim=True).clamp(min=eps).expand_as(x)


def chamfer_loss(adv, ori):  # :28-35
    return _knn_dense(adv, ori, 1)[0].squeeze(-1).mean(-1) + _knn_dense(ori, adv, 1)[0].squeeze(-1).mean(-1)


def hausdorff_loss(adv, ori):  # :45-50
    return _knn_dense(adv, ori, 1)[0].squeeze(-1).max(-1)[0]


def _kappa(pc, normal, k):
    b, _, n = pc.shape
    idx = _knn_dense(pc, pc, k + 1)[1][:, :, 1:].contiguous()
    nn_pts = torch.gather(pc, 2, idx.view(b, 1, n * k).expand(b, 3, n * k)).view(b, 3, n, k)
    vectors = _normalize(nn_pts - pc.unsqueeze(3))
    return torch.abs((vectors * normal.unsqueeze(3)).sum(1)).mean(2)


def get_kappa_ori(pc, normal, k):  # :52-62
    return _kappa(pc, normal, k)


def get_kappa_adv(adv, ori, ori_normal, k):  # :64-82
    b, _, n = adv.shape
    idx = _knn_dense(adv, ori, 1)[1]
    normal = torch.gather(ori_normal, 2, idx.view(b, 1, n).expand(b, 3, n))
    return _kappa(adv, normal, k), normal


def curvature_loss(adv, ori, adv_kappa, ori_kappa):  # :84-97
    idx = _knn_dense(adv, ori, 1)[1].squeeze(-1)
    return ((adv_kappa - torch.gather(ori_kappa, 1, idx)) ** 2).mean(-1)


def constrain_loss(adv, ori, normal, kappa_ori, k=16, w_cd=1.0, w_hd=0.1, w_curv=1.0):
    """geoA3_attack.py:131-162 (four separate kNN passes, like the reference)."""
    cd = chamfer_loss(adv, ori)
    hd = hausdorff_loss(adv, ori)
    kap, _ = get_kappa_adv(adv, ori, normal, k)
    cu = curvature_loss(adv, ori, kap, kappa_ori)
    return w_cd * cd + w_hd * hd + w_curv * cu, cd, hd, cu


class CpuAttackStep(object):
    """One reference-style attack iteration on the host CPU (net fwd, CE, losses, backward, Adam)."""

    def __init__(self, net, pc_ori, normal_ori, target, k=16, lr=0.01, initial_const=10.0, seed=0):
        self.net, self.pc, self.nrm, self.target, self.k = net, pc_ori, normal_ori, target, k
        self.kappa_ori = get_kappa_ori(pc_ori, normal_ori, k)
        g = torch.Generator().manual_seed(seed)
        self.offset = torch.empty_like(pc_ori).normal_(0.0, 1e-3, generator=g).requires_grad_(True)
        self.opt = torch.optim.Adam([self.offset], lr=lr)
        self.scale_const = torch.full((pc_ori.size(0),), float(initial_const))

    def step(self, with_net=True):
        adv = self.pc + self.offset
        if with_net:
            logits = self.net(adv)
            cls = -torch.nn.functional.cross_entropy(logits, self.target, reduction="none")
        else:
            cls = torch.zeros(adv.size(0))
        con, cd, hd, cu = constrain_loss(adv, self.pc, self.nrm, self.kappa_ori, self.k)
        loss_n = cls + self.scale_const * con
        loss = loss_n.mean()
        self.opt.zero_grad()
        loss.backward()
        self.opt.step()
        return loss_n.detach()


def time_cpu_attack(net, pc, nrm, target, steps, warmup, k=16, with_net=True):
    """-> seconds per iteration (median over `steps`), on however many threads torch currently uses."""
    st = CpuAttackStep(net, pc, nrm, target, k)
    for _ in range(warmup):
        st.step(with_net)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        st.step(with_net)
        ts.append(time.perf_counter() - t0)
    ts.sort()
    return ts[len(ts) // 2], sum(ts)
