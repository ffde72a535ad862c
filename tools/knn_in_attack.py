"""Times the neighbour-list kernels on the states a REAL attack produces (hints = the previous step's lists, clouds one
Adam step apart), at several points of the optimisation.   python tools/knn_in_attack.py [--batch 250]  -> JSON lines"""
import argparse
import json
import os
import os.path as osp
import sys

import torch

ROOT = osp.dirname(osp.dirname(osp.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from geoa3_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=250)
ap.add_argument("--k", type=int, default=16)
a = ap.parse_args()
dev = torch.device("cuda", 0)
st, pins = bench.build_state("PointNet", a.batch, bench.NPTS, 0, a.batch, dev)
k = a.k
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def t(fn, iters=5):
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return round(sorted(ts)[len(ts) // 2], 1)


probe = {1, 5, 20, 150, 499}
for step in range(500):
    if step in probe and step > 0:
        hint = st.hints.nbr[k].clone()              # lists of the previous step
        prev_adv = (st.base + st.offset).detach().clone()
    st.step()
    if step in probe and step > 0:
        adv = (st.base + st.offset).detach().contiguous()   # the cloud the NEXT step will search
        hb = st.hints
        hint = hb.nbr[k].clone()                    # = lists of `prev_adv`... of the step just done (cloud before this update)
        arr = ops.arrange(adv, hb.perm, with_bbox=True)
        move = float((adv - prev_adv).abs().max())
        new = ops.knn(adv, adv, k + 1, drop=1)[0]
        changed = float((new.sort(-1)[0] != hint.sort(-1)[0]).any(-1).float().mean())
        extra = {}
        r = dict(step=step, max_move=round(move, 5), rows_with_changed_set=round(changed, 3), **extra,
                 set_pruned=t(lambda: ops.knn(adv, adv, k + 1, drop=1, hint=hint.clone(), perm_q=hb.perm, perm_c=hb.perm,
                                              iperm_c=hb.iperm, arranged=arr, members_only=True)),
                 set_hinted=t(lambda: ops.knn(adv, adv, k + 1, drop=1, hint=hint.clone(), members_only=True)),
                 sorted_hinted=t(lambda: ops.knn(adv, adv, k + 1, drop=1, hint=hint.clone())),
                 sorted_plain=t(lambda: ops.knn(adv, adv, k + 1, drop=1)),
                 clone_only=t(lambda: hint.clone()))
        print(json.dumps(r), flush=True)
