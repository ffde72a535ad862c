"""Launches the seeded, slab-ordered nn_pair once at B=250, N=1024 (ncu target; never quote timings from this script)."""
import os.path as osp
import sys

import numpy as np
import torch

sys.path.insert(0, osp.dirname(osp.dirname(osp.abspath(__file__))))
from geoa3_b200 import ops, synth  # noqa: E402

b, n = 250, 1024
pc, nr, _ = synth.make_batch(10, n)
ori = torch.from_numpy(np.tile(pc, (25, 1, 1))).cuda()
adv = ori + torch.from_numpy(synth.make_offsets(b, n)).cuda()
d1, js, d2, is_ = ops.nn_pair(adv, ori)
perm, iperm = ops.visit_order(ori)
kw = dict(hint_a2o=js, hint_o2a=is_, perm_a=perm, perm_o=perm, iperm_a=iperm, iperm_o=iperm, ori_arranged=ops.arrange(ori, perm))
chk = ops.nn_pair(adv, ori, **kw)
torch.cuda.synchronize()
assert torch.equal(chk[1], js) and torch.equal(chk[3], is_)
print("done")
