"""Launches the kNN variants once at B=250, N=1024, k=16 (ncu target; never quote timings from this script)."""
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
adv_prev = adv - 0.003 * torch.sign(torch.randn_like(adv))
nbr_prev = ops.knn(adv_prev, adv_prev, k + 1, drop=1)[0]          # launch 0: unhinted
perm, iperm = ops.morton_order(ori)
a = ops.knn(adv, adv, k + 1, drop=1, hint=nbr_prev)[0]            # launch 1: hinted
c = ops.knn(adv, adv, k + 1, drop=1, hint=nbr_prev, perm_q=perm, perm_c=perm, iperm_c=iperm)[0]  # launch 2: hinted + pruned
sp, isp = ops.slab_order(ori)
d = ops.knn(adv, adv, k + 1, drop=1, hint=nbr_prev, perm_q=sp, perm_c=sp, iperm_c=isp)[0]       # launch 3: hinted + slab-pruned
torch.cuda.synchronize()
assert torch.equal(a, c) and torch.equal(a, d)
if "--time" in sys.argv:
    import json
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    def t(f, iters=20):
        ts = []
        for _ in range(iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); f(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        return round(sorted(ts)[len(ts) // 2], 1)
    adv_m, adv_s = ops.arrange(adv, perm), ops.arrange(adv, sp)
    d1, js, d2, is_ = ops.nn_pair(adv, ori)
    ori_m, ori_s = ops.arrange(ori, perm), ops.arrange(ori, sp)
    mk = dict(hint_a2o=js, hint_o2a=is_, perm_a=perm, perm_o=perm, iperm_a=iperm, iperm_o=iperm, ori_arranged=ori_m)
    sk = dict(hint_a2o=js, hint_o2a=is_, perm_a=sp, perm_o=sp, iperm_a=isp, iperm_o=isp, ori_arranged=ori_s)
    for kw in (mk, sk):
        chk = ops.nn_pair(adv, ori, **kw)
        assert torch.equal(chk[1], js) and torch.equal(chk[3], is_) and torch.equal(chk[0], d1)
    arr = ops.arrange(adv, sp, with_bbox=True)
    assert torch.equal(ops.knn(adv, adv, k + 1, drop=1, hint=nbr_prev, perm_q=sp, perm_c=sp, iperm_c=isp, arranged=arr)[0], a)
    res = dict(nn_pair_morton=t(lambda: ops.nn_pair(adv, ori, **mk)), nn_pair_slab=t(lambda: ops.nn_pair(adv, ori, **sk)),
               nn_pair_slab_prearranged=t(lambda: ops.nn_pair(adv, ori, adv_arranged=arr[0], **sk)),
               knn_slab_prearranged=t(lambda: ops.knn(adv, adv, k + 1, drop=1, hint=nbr_prev, perm_q=sp, perm_c=sp, iperm_c=isp, arranged=arr)),
               arrange_bbox=t(lambda: ops.arrange(adv, sp, with_bbox=True)),
               hinted=t(lambda: ops.knn(adv, adv, k + 1, drop=1, hint=nbr_prev)),
               morton_incl_gather=t(lambda: ops.knn(adv, adv, k + 1, drop=1, hint=nbr_prev, perm_q=perm, perm_c=perm, iperm_c=iperm)),
               slab_incl_gather=t(lambda: ops.knn(adv, adv, k + 1, drop=1, hint=nbr_prev, perm_q=sp, perm_c=sp, iperm_c=isp)),
               gather_only=t(lambda: ops.arrange(adv, sp)))
    print(json.dumps(res))
print("done")
