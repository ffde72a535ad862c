"""One small launch of every round-2b kernel (compute-sanitizer --tool racecheck target; b=2, n=1024, k=16)."""
import os.path as osp
import sys

import numpy as np
import torch

sys.path.insert(0, osp.dirname(osp.dirname(osp.abspath(__file__))))
from geoa3_b200 import loss_utils as L, ops, synth  # noqa: E402

b, n, k = 2, 1024, 16
pc, nr, _ = synth.make_batch(b, n)
ori, nrm = torch.from_numpy(pc).cuda(), torch.from_numpy(nr).cuda()
ko = L._get_kappa_ori(ori, nrm, k)
hb = L.HintBuffers()
for s in range(3):  # unhinted -> careful path, then hinted -> fast path
    adv = (ori + 0.01 * (s + 1) * torch.sign(torch.randn_like(ori))).requires_grad_(True)
    L.clear_cache()
    L.geo_loss(adv, ori, nrm, ko, k, 1.0, 0.1, 1.0, hints=hb)[0].sum().backward()
cells = ops.cell_sort(adv.detach(), kref=17)
ops.knn_cells(cells, k + 1, drop=1)                      # no hint: careful path
ops.knn_cells(cells, k + 1, drop=0, hint=hb.nbr[k])      # drop = 0
torch.cuda.synchronize()
print("done")
