"""Launches the neighbour-search variants once at B=250, N=1024, k=16 (ncu target; never quote timings from this script)."""
import os.path as osp
import sys

import numpy as np
import torch

sys.path.insert(0, osp.dirname(osp.dirname(osp.abspath(__file__))))
from geoa3_b200 import ops, synth  # noqa: E402

b, n, k = 250, 1024, 16
pc, nr, _ = synth.make_batch(10, n)
ori = torch.from_numpy(np.tile(pc, (25, 1, 1))).cuda()
adv = ori + torch.from_numpy(synth.make_offsets(b, n)).cuda()
step = 0.01 if "--big-step" in sys.argv else 0.003
adv_prev = adv - step * torch.sign(torch.randn_like(adv))
nbr_prev = ops.knn(adv_prev, adv_prev, k + 1, drop=1)[0]          # knn_kernel launch 0: sorted, unhinted
a = ops.knn(adv, adv, k + 1, drop=1, hint=nbr_prev)[0]            # knn_kernel launch 1: sorted, hinted
for name, fn in (("morton", ops.morton_order), ("slab", ops.slab_order)):
    p, ip = fn(ori)
    arr = ops.arrange(adv, p, with_bbox=True)
    m = ops.knn(adv, adv, k + 1, drop=1, hint=nbr_prev, perm_q=p, perm_c=p, iperm_c=ip, arranged=arr, members_only=True)[0]
    assert torch.equal(m.sort(-1)[0], a.sort(-1)[0])              # knn_members launches 0 (morton), 1 (slab)
torch.cuda.synchronize()
print("done")
