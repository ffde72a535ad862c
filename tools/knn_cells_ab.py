"""Cell-grid kNN (geoa3_cell_sort + geoa3_knn_cells) on the GPU box: timing on the states a REAL attack produces (hints =
previous step's lists) next to the slab-pruned member-set kernel, for a sweep of grids (exactness: tests/test_cells_gpu.py).
    python tools/knn_cells_ab.py [--batch 250] [--grids k8,k17,12]  -> JSON lines"""
import argparse
import json
import os.path as osp
import sys

import numpy as np
import torch

ROOT = osp.dirname(osp.dirname(osp.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, osp.join(ROOT, "tests"))
import bench  # noqa: E402
from geoa3_b200 import ops, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=250)
ap.add_argument("--k", type=int, default=16)
ap.add_argument("--skip-check", action="store_true")
ap.add_argument("--grids", type=str, default="k8,k12,k17,k24,12")
a = ap.parse_args()
dev = torch.device("cuda", 0)


def cu(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


st, pins = bench.build_state("PointNet", a.batch, bench.NPTS, 0, a.batch, dev)
k = a.k
n = bench.NPTS
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def t(fn, iters=7):
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return round(sorted(ts)[len(ts) // 2], 1)


probe = {1, 5, 20, 150, 499}
pm, ipm = ops.visit_order(st.pc_ori.detach().contiguous())
grids = a.grids.split(",")   # "k<kref>" = per-cloud adaptive grid for balls of kref points; "<g>" = fixed g^3 grid
for step in range(500):
    st.step()
    if step in probe:
        adv = (st.base + st.offset).detach().contiguous()   # the cloud the NEXT step will search
        hb = st.hints
        hint = hb.nbr[k].clone()                            # lists of the cloud before this update
        arr = ops.arrange(adv, pm, with_bbox=True)
        new = ops.knn(adv, adv, k + 1, drop=1)[0]
        out = torch.empty_like(hint)
        r = dict(step=step,
                 set_pruned=t(lambda: ops.knn(adv, adv, k + 1, drop=1, hint=hint, out=out, perm_q=pm, perm_c=pm,
                                              iperm_c=ipm, arranged=arr, members_only=True)))
        for G in grids:
            kw = dict(kref=float(G[1:])) if G[0] == "k" else dict(grid=int(G))
            cells = ops.cell_sort(adv, **kw)
            gs = G
            r["sort_" + gs] = t(lambda: ops.cell_sort(adv, out=cells, **kw))
            r["cells_" + gs] = t(lambda: ops.knn_cells(cells, k + 1, drop=1, hint=hint, out=out))
            same = torch.equal(out.sort(-1)[0], new.sort(-1)[0])
            if not same:
                r["MISMATCH_" + gs] = True
        print(json.dumps(r), flush=True)
