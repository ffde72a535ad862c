"""Cell-grid kNN (geoa3_cell_sort + geoa3_knn_cells) on the GPU box: exactness against the C oracle on the case list of
tests/test_loss_gpu.py::test_knn_set_members_exact, then timing on the states a REAL attack produces (hints = previous
step's lists) next to the slab-pruned member-set kernel, for a sweep of grid sizes.
    python tools/knn_cells_ab.py [--batch 250] [--skip-check]  -> JSON lines"""
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
ap.add_argument("--grids", type=str, default="6,7,8,9,10,12")
a = ap.parse_args()
dev = torch.device("cuda", 0)


def cu(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


if not a.skip_check:
    from oracle import oracle as O

    def make(b, n, seed, std):
        pc, nr, _ = synth.make_batch(b, n, seed)
        return (pc + synth.make_offsets(b, n, seed=seed + 100, std=std)).astype(np.float32), nr

    rng = np.random.default_rng(5)
    lat, _ = synth.lattice_cloud(343)
    dup = make(2, 300, 4, 1e-2)[0]
    dup[:, :, 150:200] = dup[:, :, 0:50]
    shifted = make(2, 500, 7, 1e-2)[0] + np.float32(100.0)
    tiny = make(2, 500, 8, 1e-2)[0] * np.float32(1e-3)
    flat = make(2, 400, 9, 1e-2)[0]
    flat[:, 2, :] = 0.25
    same = np.zeros((1, 3, 64), np.float32) + np.float32(0.5)
    cases = [(make(3, 1024, 2, 2e-2)[0], 17), (make(2, 1024, 1, 1e-2)[0], 33), (make(2, 777, 5, 5e-2)[0], 9),
             (np.stack([lat, lat * 0.5]), 9), (np.ascontiguousarray(dup), 17), (make(1, 2500, 3, 1e-2)[0], 17),
             (make(2, 40, 0, 1e-1)[0], 33), (make(2, 20, 0, 1e-1)[0], 17), (shifted, 17), (tiny, 17), (flat, 17),
             (same, 9), (make(1, 4096, 11, 1e-2)[0], 17), (make(1, 10000, 12, 5e-3)[0], 17), (make(2, 1024, 13, 1e-2)[0], 11)]
    bad = 0
    for ci, (pts, K) in enumerate(cases):
        b, _, n = pts.shape
        K = min(K, n)
        P = cu(pts)
        oi, od = O.knn(pts, pts, K)
        want_i, want_d = oi[:, :, 1:], od[:, :, 1:]
        exact = cu(want_i)
        stale = cu(O.knn(pts + 0.05, pts[:, :, ::-1].copy(), K)[0][:, :, 1:])
        junk = cu(rng.integers(-3, n + 50, (b, n, K - 1)).astype(np.int32))
        zeros = torch.zeros(b, n, K - 1, dtype=torch.int32, device="cuda")
        gmax = ops._lib.load().geoa3_cell_grid_max(n)
        base = None
        for G in sorted({1, 2, 5, ops.cell_grid_size(n, K), min(gmax, 13)}):
            blobs = ops.cell_sort(P, G)
            first = None
            for h, tag in ((None, "none"), (exact, "exact"), (stale, "stale"), (junk, "junk"), (zeros, "dup")):
                idx, dist = ops.knn_cells(blobs, n, G, K, drop=1, return_dist=True, hint=h)
                i_, d_ = idx.cpu().numpy(), dist.cpu().numpy()
                key = d_.view(np.int32).astype(np.int64) * 65536 + i_
                o = np.argsort(key, -1)
                ok = (np.array_equal(np.take_along_axis(i_, o, -1), want_i)
                      and np.array_equal(np.take_along_axis(d_, o, -1), want_d))
                if first is None:
                    first = idx
                ok_order = torch.equal(idx, first)
                if not (ok and ok_order):
                    bad += 1
                    wrong = int((np.take_along_axis(i_, o, -1) != want_i).any(-1).sum())
                    print(json.dumps(dict(case=ci, n=n, K=K, G=G, hint=tag, members_ok=bool(ok), order_ok=bool(ok_order),
                                          wrong_rows=wrong)), flush=True)
        # in place: the buffer is hint and output at once
        buf, cur = exact.clone(), pts
        G = ops.cell_grid_size(n, K)
        for step in range(3):
            cur = (cur + synth.make_offsets(b, n, seed=40 + step, std=3e-3)).astype(np.float32)
            C = cu(cur)
            ops.knn_cells(ops.cell_sort(C, G), n, G, K, drop=1, hint=buf, out=buf)
            if not np.array_equal(np.sort(buf.cpu().numpy(), -1), np.sort(O.knn(cur, cur, K)[0][:, :, 1:], -1)):
                bad += 1
                print(json.dumps(dict(case=ci, inplace_step=step, ok=False)), flush=True)
    print(json.dumps(dict(check="knn_cells vs oracle", cases=len(cases), failures=bad)), flush=True)

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
grids = [int(g) for g in a.grids.split(",")]
for step in range(500):
    st.step()
    if step in probe:
        adv = (st.base + st.offset).detach().contiguous()   # the cloud the NEXT step will search
        hb = st.hints
        hint = hb.nbr[k].clone()                            # lists of the cloud before this update
        arr = ops.arrange(adv, hb.perm, with_bbox=True)
        new = ops.knn(adv, adv, k + 1, drop=1)[0]
        out = torch.empty_like(hint)
        r = dict(step=step,
                 set_pruned=t(lambda: ops.knn(adv, adv, k + 1, drop=1, hint=hint, out=out, perm_q=hb.perm, perm_c=hb.perm,
                                              iperm_c=hb.iperm, arranged=arr, members_only=True)))
        for G in grids:
            blobs = ops.cell_sort(adv, G)
            r["sort_G%d" % G] = t(lambda: ops.cell_sort(adv, G, out=blobs))
            r["cells_G%d" % G] = t(lambda: ops.knn_cells(blobs, n, G, k + 1, drop=1, hint=hint, out=out))
            same = torch.equal(out.sort(-1)[0], new.sort(-1)[0])
            if not same:
                r["MISMATCH_G%d" % G] = True
        print(json.dumps(r), flush=True)
