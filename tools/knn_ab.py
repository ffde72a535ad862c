"""A/B of the two exact kNN kernels (run on the GPU box): geoa3_knn_set (knn_select.cu: members in visiting order)
against geoa3_knn (knn.cu: sorted by distance) — the member sets and their distances must agree bit for bit — plus
CUDA-event timings.   python tools/knn_ab.py [--quick]   -> JSON lines  ([set_us, sorted_us, same])"""
import json
import os
import os.path as osp
import sys

import numpy as np
import torch

sys.path.insert(0, osp.dirname(osp.dirname(osp.abspath(__file__))))
from geoa3_b200 import ops, synth  # noqa: E402

FLUSH = None


def timeit(fn, iters=10, warm=3):
    global FLUSH
    if FLUSH is None:
        FLUSH = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        FLUSH.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return round(ts[len(ts) // 2], 1)


def canon(idx, dist):
    """rows sorted by (distance, index) — the order geoa3_knn writes"""
    key = (dist.double() * 0 + dist.view(torch.int32).long()) * 65536 + idx.long()
    o = key.argsort(-1)
    return idx.gather(-1, o), dist.gather(-1, o)


def both(fn):
    """fn(members_only) -> (idx, dist)"""
    a = fn(True)
    ta = timeit(lambda: fn(True))
    b = fn(False)
    tb = timeit(lambda: fn(False))
    ca = canon(*a)
    same = torch.equal(ca[0], b[0]) and torch.equal(ca[1], b[1])
    return ta, tb, same


def case(b, n, k, std=1e-2, step=0.003):
    base = min(b, 16)
    pc, _, _ = synth.make_batch(base, n)
    reps = (b + base - 1) // base
    ori = torch.from_numpy(np.tile(pc, (reps, 1, 1))[:b].copy()).cuda()
    adv = ori + torch.from_numpy(synth.make_offsets(b, n, std=std)).cuda()
    prev = (adv - step * torch.sign(torch.randn_like(adv))).contiguous()
    hn = ops.knn(prev, prev, k + 1, drop=1)[0]
    pm, ipm = ops.visit_order(ori)
    r = dict(b=b, n=n, k=k)
    arr = ops.arrange(adv, pm, with_bbox=True)
    r["plain"] = both(lambda mo: ops.knn(adv, adv, k + 1, drop=1, return_dist=True, members_only=mo))
    r["hinted"] = both(lambda mo: ops.knn(adv, adv, k + 1, drop=1, hint=hn, return_dist=True, members_only=mo))
    r["pruned_incl_arrange"] = both(lambda mo: ops.knn(adv, adv, k + 1, drop=1, hint=hn, perm_q=pm, perm_c=pm, iperm_c=ipm,
                                                       return_dist=True, members_only=mo))
    r["pruned_prearranged"] = both(lambda mo: ops.knn(adv, adv, k + 1, drop=1, hint=hn, perm_q=pm, perm_c=pm, iperm_c=ipm,
                                                      arranged=arr, return_dist=True, members_only=mo))
    # stale hints (a different cloud's neighbours): exact all the same, only slower
    stale = hn.roll(1, 0).contiguous()
    r["stale_hint"] = both(lambda mo: ops.knn(adv, adv, k + 1, drop=1, hint=stale, return_dist=True, members_only=mo))
    # in-place refresh of persistent hint buffers (what the attack does), 3 consecutive steps
    buf = hn.clone()
    cur = adv
    ok = True
    for _ in range(3):
        cur = (cur + step * torch.sign(torch.randn_like(cur))).contiguous()
        ops.knn(cur, cur, k + 1, drop=1, hint=buf, out=buf, perm_q=pm, perm_c=pm, iperm_c=ipm, members_only=True)
        ref = ops.knn(cur, cur, k + 1, drop=1)[0]
        ok &= torch.equal(buf.sort(-1)[0], ref.sort(-1)[0])
        # order independent of the hint: an unhinted search writes the same rows
        ok &= torch.equal(buf, ops.knn(cur, cur, k + 1, drop=1, perm_q=pm, perm_c=pm, iperm_c=ipm, members_only=True)[0])
    r["inplace_3steps_exact"] = bool(ok)
    print(json.dumps(r), flush=True)


def edge_cases():
    torch.manual_seed(1)
    res = {}
    for (b, n, m, K, drop) in ((3, 12, 12, 12, 0), (2, 40, 7, 5, 1), (4, 300, 300, 33, 1), (2, 2500, 2500, 17, 1),
                               (2, 100, 3000, 9, 0), (1, 5000, 64, 3, 0)):
        q = torch.randn(b, 3, n, device="cuda")
        c = q if n == m else torch.randn(b, 3, m, device="cuda")
        # duplicates + a lattice: exact ties on distance must resolve on the index
        c = (c * 4).round() / 4
        q = c if n == m else (q * 4).round() / 4
        _, _, same = both(lambda mo: ops.knn(q, c, K, drop=drop, return_dist=True, members_only=mo))
        res["b%d_n%d_m%d_K%d" % (b, n, m, K)] = same
    # far-from-origin cloud: the filter margin scales with |p|^2, results stay exact
    q = torch.randn(2, 3, 1024, device="cuda") * 0.01 + 300.0
    _, _, same = both(lambda mo: ops.knn(q, q, 17, drop=1, return_dist=True, members_only=mo))
    res["shifted_cloud"] = same
    q = torch.randn(2, 3, 1024, device="cuda") * 1e-4
    _, _, same = both(lambda mo: ops.knn(q, q, 17, drop=1, return_dist=True, members_only=mo))
    res["tiny_cloud"] = same
    print(json.dumps(dict(what="edge_cases", **res)), flush=True)


if __name__ == "__main__":
    edge_cases()
    case(250, 1024, 16)
    if "--quick" not in sys.argv:
        case(250, 1024, 32)
        case(32, 1024, 16)
        case(64, 4096, 16)
        case(64, 4096, 32)
        case(64, 10000, 16)
        case(64, 10000, 32)
        case(250, 512, 8)
