"""Per-kernel CUDA-event timings of the hot path at BASELINE config sizes (run on the GPU box).
   python tools/time_kernels.py [--b 250 --n 1024 --k 16]   -> JSON lines on stdout"""
import argparse
import json
import os.path as osp
import sys

import numpy as np
import torch

sys.path.insert(0, osp.dirname(osp.dirname(osp.abspath(__file__))))
from geoa3_b200 import ops  # noqa: E402
from geoa3_b200 import synth  # noqa: E402


def timeit(fn, iters=20, warm=5, flush=None):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def sweep():
    """kNN + kappa kernels at B=64, N in {1024,4096,10000}, k in {16,32} (BASELINE config[3])."""
    B = 64
    for n in (1024, 4096, 10000):
        pc, nr, _ = synth.make_batch(16, n)
        ori = torch.from_numpy(np.tile(pc, (4, 1, 1))).cuda()
        nrm = torch.from_numpy(np.tile(nr, (4, 1, 1))).cuda()
        adv = ori + torch.from_numpy(synth.make_offsets(B, n)).cuda()
        prev = (adv - 0.003 * torch.sign(torch.randn_like(adv))).contiguous()
        for k in (16, 32):
            hn = ops.knn(prev, prev, k + 1, drop=1)[0]
            t_knn = timeit(lambda: ops.knn(adv, adv, k + 1, drop=1), iters=5, warm=2)
            t_hint = timeit(lambda: ops.knn(adv, adv, k + 1, drop=1, hint=hn), iters=5, warm=2)
            pm, ipm = ops.morton_order(ori)
            t_mort = timeit(lambda: ops.knn(adv, adv, k + 1, drop=1, hint=hn, perm_q=pm, perm_c=pm, iperm_c=ipm), iters=5, warm=2)
            assert torch.equal(ops.knn(adv, adv, k + 1, drop=1, hint=hn, perm_q=pm, perm_c=pm, iperm_c=ipm)[0],
                               ops.knn(adv, adv, k + 1, drop=1)[0])
            nbr = ops.knn(adv, adv, k + 1, drop=1, hint=hn)[0]
            assert torch.equal(nbr, ops.knn(adv, adv, k + 1, drop=1)[0])
            _, js, _, _ = ops.nn_pair(adv, ori)
            t_kap = timeit(lambda: ops.kappa_loss_fwd(adv, normal=nrm, jstar=js, nbr=nbr), iters=5, warm=2)
            byts = (28 + 4 * k) * B * n
            flop = 8.0 * B * n * n
            print(json.dumps(dict(what="curvature_sweep", B=B, n=n, k=k, knn_us=round(t_knn[0], 1),
                                  knn_hinted_us=round(t_hint[0], 1), knn_hinted_morton_us=round(t_mort[0], 1),
                                  kappa_us=round(t_kap[0], 1),
                                  hbm_frac=round(byts / ((t_hint[0] + t_kap[0]) * 1e-6) / 6555.2e9, 5),
                                  fp32_tflops_hinted=round(flop / (t_hint[0] * 1e-6) / 1e12, 2))))


def fps_plain():
    """farthest_points_sample (Lib/utility.py:175-187): one launch here vs the reference's loop of num_points-1 torch
    passes (restated with plain torch ops and run on the same GPU)."""
    from geoa3_b200 import utility

    def torch_loop(pts, m, start):
        b, _, n = pts.shape
        sel = start.long()[:, None]
        dists = torch.full((b, n), float("inf"), device=pts.device)
        for _ in range(m - 1):
            last = torch.gather(pts, 2, sel[:, -1][:, None, None].expand(b, 3, 1))
            dists = torch.min(dists, torch.norm(pts - last, dim=1))
            sel = torch.cat([sel, torch.argmax(dists, dim=1, keepdim=True)], dim=1)
        return torch.gather(pts, 2, sel[:, None, :].expand(b, 3, m))

    for (b, n, m) in ((64, 4096, 1024), (64, 10000, 1024), (250, 2048, 1024)):
        pc, _, _ = synth.make_batch(8, n)
        pts = torch.from_numpy(np.tile(pc, ((b + 7) // 8, 1, 1))[:b].copy()).cuda()
        start = torch.randint(n, (b,), device="cuda", dtype=torch.int32)
        ours = utility.farthest_points_sample(pts, m, start=start)
        ref = torch_loop(pts, m, start)
        t_o = timeit(lambda: utility.farthest_points_sample(pts, m, start=start), iters=5, warm=2)
        t_r = timeit(lambda: torch_loop(pts, m, start), iters=2, warm=1)
        print(json.dumps(dict(what="farthest_points_sample", b=b, n=n, m=m, ours_us=round(t_o[0], 1),
                              torch_loop_us=round(t_r[0], 1), same_picks=bool(torch.equal(ours, ref)))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--b", type=int, default=250)
    ap.add_argument("--n", type=int, default=1024)
    ap.add_argument("--k", type=int, default=16)
    ap.add_argument("--pn2", action="store_true")
    ap.add_argument("--sweep", action="store_true", help="BASELINE config[3]: curvature-loss scaling sweep")
    ap.add_argument("--fps-plain", dest="fps_plain", action="store_true", help="farthest_points_sample vs the torch loop")
    a = ap.parse_args()
    if a.sweep:
        return sweep()
    if a.fps_plain:
        return fps_plain()
    b, n, k = a.b, a.n, a.k
    pc, nr, _ = synth.make_batch(min(b, 20), n)
    reps = (b + pc.shape[0] - 1) // pc.shape[0]
    ori = torch.from_numpy(np.tile(pc, (reps, 1, 1))[:b].copy()).cuda()
    nrm = torch.from_numpy(np.tile(nr, (reps, 1, 1))[:b].copy()).cuda()
    adv = ori + torch.from_numpy(synth.make_offsets(b, n)).cuda()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    res = {}
    res["nn_pair"] = timeit(lambda: ops.nn_pair(adv, ori), flush=flush)
    d1, js, d2, is_ = ops.nn_pair(adv, ori)
    res["nn_pair_hinted"] = timeit(lambda: ops.nn_pair(adv, ori, hint_a2o=js, hint_o2a=is_), flush=flush)
    perm, iperm = ops.visit_order(ori)
    ori_s = ops.arrange(ori, perm)
    mk = dict(hint_a2o=js, hint_o2a=is_, perm_a=perm, perm_o=perm, iperm_a=iperm, iperm_o=iperm, ori_arranged=ori_s)
    res["nn_pair_ordered(incl. arrange of adv)"] = timeit(lambda: ops.nn_pair(adv, ori, **mk), flush=flush)
    chk = ops.nn_pair(adv, ori, **mk)
    assert torch.equal(chk[1], js) and torch.equal(chk[3], is_) and torch.equal(chk[0], d1)
    perm, iperm = ops.morton_order(ori)
    res["knn_self"] = timeit(lambda: ops.knn(adv, adv, k + 1, drop=1), flush=flush)
    nbr = ops.knn(adv, adv, k + 1, drop=1)[0]
    # hint = neighbours of the previous attack step (adv moved by one Adam step ~ lr*sign = 0.01 per coordinate at most)
    adv_prev = adv - 0.003 * torch.sign(torch.randn_like(adv))
    nbr_prev = ops.knn(adv_prev, adv_prev, k + 1, drop=1)[0]
    res["knn_self_hinted"] = timeit(lambda: ops.knn(adv, adv, k + 1, drop=1, hint=nbr_prev), flush=flush)
    assert torch.equal(ops.knn(adv, adv, k + 1, drop=1, hint=nbr_prev)[0], nbr)
    perm, iperm = ops.morton_order(ori)
    res["knn_self_morton(incl. gather)"] = timeit(lambda: ops.knn(adv, adv, k + 1, drop=1, hint=nbr_prev, perm_q=perm, perm_c=perm,
                                                    iperm_c=iperm), flush=flush)
    assert torch.equal(ops.knn(adv, adv, k + 1, drop=1, hint=nbr_prev, perm_q=perm, perm_c=perm, iperm_c=iperm)[0], nbr)
    nbr_o = ops.knn(ori, ori, k + 1, drop=1)[0]
    ko = ops.kappa_loss_fwd(ori, normal=nrm, nbr=nbr_o)["kappa"]
    f = lambda: ops.kappa_loss_fwd(adv, normal=nrm, jstar=js, nbr=nbr, d_a2o=d1, d_o2a=d2, kappa_ori=ko, want_nrm=True,
                                   want_cd=True, want_hd=True, want_curv=True)
    res["kappa_loss_fwd"] = timeit(f, flush=flush)
    out = f()
    g = torch.full((b,), 1.0 / b, device="cuda")
    res["loss_bwd"] = timeit(lambda: ops.loss_bwd(adv, ori=ori, nrm_adv=out["nrm"], kappa_adv=out["kappa"], kappa_ori=ko,
                                                  jstar=js, istar=is_, nbr=nbr, hd_arg=out["hd_arg"], g_cd=g, g_hd=g,
                                                  g_cu=g), flush=flush)
    tot = res["nn_pair_ordered(incl. arrange of adv)"][0] + res["knn_self_hinted"][0] + res["kappa_loss_fwd"][0] + res["loss_bwd"][0]
    pairs = b * n * n
    print(json.dumps(dict(what="loss_path", b=b, n=n, k=k, us_median={k_: round(v[0], 2) for k_, v in res.items()},
                          us_min={k_: round(v[1], 2) for k_, v in res.items()}, total_us=round(tot, 2),
                          nn_pair_Gpairs_s=round(2 * pairs / res["nn_pair"][0] / 1e3, 1),
                          knn_Gpairs_s=round(pairs / res["knn_self"][0] / 1e3, 1),
                          hbm_frac=round(52 * b * n / (tot * 1e-6) / 6555.2e9, 4))))
    if a.pn2:
        xyz = ori.transpose(1, 2).contiguous()
        r = {}
        r["fps_1024_512"] = timeit(lambda: ops.furthest_point_sampling(xyz, 512), iters=5, warm=2)
        fi = ops.furthest_point_sampling(xyz, 512)
        new = ops.gather_points(ori, fi).transpose(1, 2).contiguous()
        r["ball_query_r.2_ns64"] = timeit(lambda: ops.ball_query(new, xyz, 0.2, 64), flush=flush)
        idx = ops.ball_query(new, xyz, 0.2, 64)
        r["group_c3"] = timeit(lambda: ops.group_points(ori, idx), flush=flush)
        feats = torch.randn(b, 128, 512, device="cuda")
        fi2 = ops.furthest_point_sampling(new, 128)
        new2 = ops.gather_points(new.transpose(1, 2).contiguous(), fi2).transpose(1, 2).contiguous()
        r["fps_512_128"] = timeit(lambda: ops.furthest_point_sampling(new, 128), iters=5, warm=2)
        idx2 = ops.ball_query(new2, new, 0.4, 64)
        r["ball_query_r.4_ns64_n512"] = timeit(lambda: ops.ball_query(new2, new, 0.4, 64), flush=flush)
        r["group_c128"] = timeit(lambda: ops.group_points(feats, idx2), flush=flush)
        go = torch.randn(b, 128, 128, 64, device="cuda")
        r["group_grad_c128"] = timeit(lambda: ops.group_points_grad(go, idx2, 512), flush=flush)
        go3 = torch.randn(b, 3, 512, 64, device="cuda")
        r["group_grad_c3"] = timeit(lambda: ops.group_points_grad(go3, idx, 1024), flush=flush)
        gb = b * 128 * 128 * 64 * 4 / 1e9
        print(json.dumps(dict(what="pointnet2", b=b, us_median={k_: round(v[0], 2) for k_, v in r.items()},
                              group_c128_GBs=round(gb / (r["group_c128"][0] * 1e-6), 1),
                              group_grad_c128_GBs=round(gb / (r["group_grad_c128"][0] * 1e-6), 1))))
        ext = None
        try:
            from oracle import build_ref
            ext = build_ref.load_ref()
        except Exception as e:  # noqa
            print(json.dumps(dict(ref_ext_error=str(e))))
        if ext is not None:
            rr = {}
            rr["fps_1024_512"] = timeit(lambda: ext.furthest_point_sampling(xyz, 512), iters=5, warm=2)
            rr["ball_query_r.2_ns64"] = timeit(lambda: ext.ball_query(new, xyz, 0.2, 64), iters=5, warm=2)
            rr["group_c3"] = timeit(lambda: ext.group_points(ori, idx), iters=5, warm=2)
            rr["group_c128"] = timeit(lambda: ext.group_points(feats, idx2), iters=5, warm=2)
            rr["group_grad_c128"] = timeit(lambda: ext.group_points_grad(go, idx2, 512), iters=5, warm=2)
            rr["fps_512_128"] = timeit(lambda: ext.furthest_point_sampling(new, 128), iters=5, warm=2)
            print(json.dumps(dict(what="pointnet2_reference_kernels_sm100", b=b,
                                  us_median={k_: round(v[0], 2) for k_, v in rr.items()})))


if __name__ == "__main__":
    main()
