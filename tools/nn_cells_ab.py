"""Cell-grid 1-NN (geoa3_nn_pair_cells) on the GPU box: timing on the states a REAL attack produces next to the
box-pruned nn_pair launch (arrange included), for a sweep of grids (exactness: tests/test_cells_gpu.py).
    python tools/nn_cells_ab.py [--batch 250] [--grids k2,k4,16] [--adv k17]  -> JSON lines"""
import argparse
import json
import os.path as osp
import sys

import numpy as np
import torch

ROOT = osp.dirname(osp.dirname(osp.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from geoa3_b200 import ops, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=250)
ap.add_argument("--skip-check", action="store_true")
ap.add_argument("--adv", type=str, default="k17")
ap.add_argument("--grids", type=str, default="k2,k4,k8,k17,16")
a = ap.parse_args()
dev = torch.device("cuda", 0)


def cu(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


st, pins = bench.build_state("PointNet", a.batch, bench.NPTS, 0, a.batch, dev)
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
grids = a.grids.split(",")   # grid of the ORIGINAL cloud: "k<kref>" adaptive, "<g>" fixed; adv uses --adv
ori = st.pc_ori.detach().contiguous()
pm, ipm = ops.visit_order(ori)
ori_arr = ops.arrange(ori, pm)
for step in range(500):
    st.step()
    if step in probe:
        adv = (st.base + st.offset).detach().contiguous()
        hb = st.hints
        hj, hi = hb.jstar.clone(), hb.istar.clone()
        outs = (torch.empty_like(hb.d1), torch.empty_like(hj), torch.empty_like(hb.d2), torch.empty_like(hi))

        def pruned():
            arr = ops.arrange(adv, pm, with_bbox=True)
            ops.nn_pair(adv, ori, hint_a2o=hj, hint_o2a=hi, perm_a=pm, perm_o=pm, iperm_a=ipm,
                        iperm_o=ipm, ori_arranged=ori_arr, adv_arranged=arr[0], out=outs)

        r = dict(step=step, nn_pair_pruned_incl_arrange=t(pruned))
        ref = [x.clone() for x in outs]
        for G in grids:
            gs = G
            kw = lambda g: dict(kref=float(g[1:])) if g[0] == "k" else dict(grid=int(g))
            bo = ops.cell_sort(ori, **kw(G))
            ba = ops.cell_sort(adv, **kw(a.adv))
            r["cells_" + gs] = t(lambda: ops.nn_pair_cells(ba, bo, hint_a2o=hj, hint_o2a=hi, out=outs))
            if not all(torch.equal(x, y) for x, y in zip(outs, ref)):
                r["MISMATCH_" + gs] = True
        print(json.dumps(r), flush=True)
