"""Cell-grid 1-NN (geoa3_nn_pair_cells) on the GPU box: bit-exactness against the C oracle (distances and argmins, both
directions, every kind of seed, several grids, ragged / shifted / tiny / duplicated clouds), then timing on the states
a REAL attack produces next to the box-pruned nn_pair launch (arrange included).
    python tools/nn_cells_ab.py [--batch 250] [--skip-check]  -> JSON lines"""
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
ap.add_argument("--grids", type=str, default="8,10,12,16")
a = ap.parse_args()
dev = torch.device("cuda", 0)


def cu(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


if not a.skip_check:
    from oracle import oracle as O

    def make(b, n, seed, std):
        pc, nr, _ = synth.make_batch(b, n, seed)
        return (pc + synth.make_offsets(b, n, seed=seed + 100, std=std)).astype(np.float32), pc

    rng = np.random.default_rng(7)
    lat, _ = synth.lattice_cloud(343)
    cases = [make(3, 1024, 2, 2e-2), make(2, 1000, 1, 5e-2), make(2, 777, 5, 1e-1),
             (make(2, 333, 2, 1e-1)[0], make(2, 777, 6, 1e-2)[1]),                     # n != m
             (np.stack([lat, lat * 0.5]), np.stack([lat * 0.5, lat])),                  # lattice ties
             tuple(x + np.float32(100.0) for x in make(2, 500, 7, 1e-2)),              # shifted
             tuple(x * np.float32(1e-3) for x in make(2, 500, 8, 1e-2)),               # tiny
             (make(2, 400, 9, 1e-2)[0] + np.float32(3.0), make(2, 400, 9, 1e-2)[1]),   # disjoint boxes
             (np.zeros((1, 3, 64), np.float32) + np.float32(0.5), make(1, 90, 3, 1e-2)[1]),
             make(1, 4096, 11, 1e-2), make(1, 10000, 12, 5e-3), make(2, 20, 0, 1e-1)]
    d = make(2, 300, 4, 1e-2)
    d[1][:, :, 150:200] = d[1][:, :, 0:50]                                            # duplicated candidates
    cases.append(d)
    bad = 0
    for ci, (adv, ori) in enumerate(cases):
        adv, ori = np.ascontiguousarray(adv, np.float32), np.ascontiguousarray(ori, np.float32)
        b, _, n = adv.shape
        m = ori.shape[2]
        od1, oj1 = O.nn1(adv, ori)
        od2, oi2 = O.nn1(ori, adv)
        A, Oc = cu(adv), cu(ori)
        for ga, go in ((1, 1), (3, 5), (ops.cell_grid_size(n, 17), ops.cell_grid_size(m, 17)), (13, 11)):
            ga = min(ga, ops._lib.load().geoa3_cell_grid_max(n))
            go = min(go, ops._lib.load().geoa3_cell_grid_max(m))
            ba, bo = ops.cell_sort(A, ga), ops.cell_sort(Oc, go)
            if max(ba.shape[1], bo.shape[1]) > 226 * 1024:
                continue
            hints = [(None, None, "none"), (cu(oj1), cu(oi2), "exact"),
                     (cu(rng.integers(-3, m + 50, (b, n)).astype(np.int32)), cu(rng.integers(-3, n + 50, (b, m)).astype(np.int32)), "junk"),
                     (torch.zeros(b, n, dtype=torch.int32, device="cuda"), torch.zeros(b, m, dtype=torch.int32, device="cuda"), "zeros")]
            for h1, h2, tag in hints:
                d1, j1, d2, i2 = ops.nn_pair_cells(ba, bo, n, m, ga, go, hint_a2o=h1, hint_o2a=h2)
                ok = (np.array_equal(j1.cpu().numpy(), oj1) and np.array_equal(i2.cpu().numpy(), oi2)
                      and np.array_equal(d1.cpu().numpy(), od1) and np.array_equal(d2.cpu().numpy(), od2))
                if not ok:
                    bad += 1
                    print(json.dumps(dict(case=ci, n=n, m=m, ga=ga, go=go, hint=tag,
                                          wrong_j=int((j1.cpu().numpy() != oj1).sum()), wrong_i=int((i2.cpu().numpy() != oi2).sum()))), flush=True)
            d1, j1, _, _ = ops.nn_pair_cells(ba, bo, n, m, ga, go, both=False)
            if not (np.array_equal(j1.cpu().numpy(), oj1) and np.array_equal(d1.cpu().numpy(), od1)):
                bad += 1
                print(json.dumps(dict(case=ci, one_sided=False)), flush=True)
    print(json.dumps(dict(check="nn_pair_cells vs oracle", cases=len(cases), failures=bad)), flush=True)

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
grids = [int(g) for g in a.grids.split(",")]
ori = st.pc_ori.detach().contiguous() if hasattr(st, "pc_ori") else st.base.detach().contiguous()
for step in range(500):
    st.step()
    if step in probe:
        adv = (st.base + st.offset).detach().contiguous()
        hb = st.hints
        hj, hi = hb.jstar.clone(), hb.istar.clone()
        outs = (torch.empty_like(hb.d1), torch.empty_like(hj), torch.empty_like(hb.d2), torch.empty_like(hi))

        def pruned():
            arr = ops.arrange(adv, hb.perm, with_bbox=True)
            ops.nn_pair(adv, ori, hint_a2o=hj, hint_o2a=hi, perm_a=hb.perm, perm_o=hb.perm, iperm_a=hb.iperm,
                        iperm_o=hb.iperm, ori_arranged=hb.ori_arranged, adv_arranged=arr[0], out=outs)

        r = dict(step=step, nn_pair_pruned_incl_arrange=t(pruned))
        ref = [x.clone() for x in outs]
        for G in grids:
            bo = ops.cell_sort(ori, G)
            ba = ops.cell_sort(adv, G)
            r["cells_G%d" % G] = t(lambda: ops.nn_pair_cells(ba, bo, n, n, G, G, hint_a2o=hj, hint_o2a=hi, out=outs))
            if not all(torch.equal(x, y) for x, y in zip(outs, ref)):
                r["MISMATCH_G%d" % G] = True
        print(json.dumps(r), flush=True)
