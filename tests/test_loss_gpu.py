"""GPU parity (-m gpu): the CUDA loss path, called through the C ABI (geoa3_b200.ops / loss_utils),
against the CPU oracle on the same seeded inputs and against the committed reference golden vectors.
Indices bit-exact; values / gradients within 1e-5 relative (north_star)."""
import numpy as np
import pytest
import torch

from helpers import golden_files, rel_err
from oracle import oracle as O
from oracle import synth

pytestmark = pytest.mark.gpu
TOL = 1e-5


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def make(b, n, start=0, std=1e-2, seed=1):
    pc, nr, _ = synth.make_batch(b, n, start)
    adv = pc + synth.make_offsets(b, n, seed=seed, std=std)
    return adv, pc, nr


@pytest.mark.parametrize("b,n,m,std", [(4, 1024, 1024, 1e-3), (3, 1000, 1000, 5e-2), (2, 333, 777, 1e-1),
                                       (1, 2500, 2049, 1e-2), (5, 7, 3, 1.0)])
def test_nn_pair_bitexact(b, n, m, std):
    from geoa3_b200 import ops

    adv, _, _ = make(b, n, 0, std)
    _, ori, _ = make(b, m, 3, std)
    d1, j1, d2, i2 = ops.nn_pair(cu(adv), cu(ori))
    od1, oj1 = O.nn1(adv, ori)
    od2, oi2 = O.nn1(ori, adv)
    assert np.array_equal(j1.cpu().numpy(), oj1) and np.array_equal(i2.cpu().numpy(), oi2)
    assert np.array_equal(d1.cpu().numpy(), od1) and np.array_equal(d2.cpu().numpy(), od2)  # bit-exact fp32
    d1b, j1b, none_d, none_i = ops.nn_pair(cu(adv), cu(ori), both=False)
    assert none_d is None and np.array_equal(j1b.cpu().numpy(), oj1)


def test_nn_pair_ties_lowest_index():
    from geoa3_b200 import ops

    pc, _ = synth.lattice_cloud(216)
    a = np.stack([pc, pc[:, ::-1].copy()])
    o = np.stack([pc + 0.125, pc])  # half-cell shift: 8 equidistant corners per query
    d1, j1, d2, i2 = ops.nn_pair(cu(a), cu(o))
    od1, oj1 = O.nn1(a, o)
    od2, oi2 = O.nn1(o, a)
    assert np.array_equal(j1.cpu().numpy(), oj1) and np.array_equal(i2.cpu().numpy(), oi2)
    assert np.array_equal(d1.cpu().numpy(), od1)


@pytest.mark.parametrize("b,n,K", [(3, 1024, 17), (2, 1024, 33), (2, 1000, 17), (2, 515, 3), (1, 4096, 17),
                                   (2, 300, 6), (1, 40, 33)])
def test_knn_bitexact(b, n, K):
    from geoa3_b200 import ops

    adv, _, _ = make(b, n, 2, 2e-2)
    idx, dist = ops.knn(cu(adv), cu(adv), K, drop=0, return_dist=True)
    oi, od = O.knn(adv, adv, K)
    assert np.array_equal(idx.cpu().numpy(), oi)
    assert np.array_equal(dist.cpu().numpy(), od)
    idx1, _ = ops.knn(cu(adv), cu(adv), K, drop=1)
    assert np.array_equal(idx1.cpu().numpy(), oi[:, :, 1:])


def test_knn_ties_lexicographic():
    from geoa3_b200 import ops

    pc, _ = synth.lattice_cloud(343)
    pc = np.stack([pc, pc * 0.5])
    for K in (9, 17, 33):
        idx, dist = ops.knn(cu(pc), cu(pc), K, return_dist=True)
        oi, od = O.knn(pc, pc, K)
        assert np.array_equal(idx.cpu().numpy(), oi), K
    # duplicated points: zero distances, self is not necessarily column 0
    dup = np.concatenate([pc[:, :, :50], pc[:, :, :50], pc[:, :, 50:100]], 2)
    idx, _ = ops.knn(cu(dup), cu(dup), 5)
    assert np.array_equal(idx.cpu().numpy(), O.knn(dup, dup, 5)[0])


@pytest.mark.parametrize("b,n,k,std", [(4, 1024, 16, 1e-3), (3, 1000, 16, 3e-2), (2, 512, 32, 1e-2), (2, 200, 2, 1e-1)])
def test_fused_forward_backward_vs_oracle(b, n, k, std):
    from geoa3_b200 import loss_utils as L

    adv, ori, nrm = make(b, n, 5, std)
    kap_ori_o, nbr_o = O.kappa_ori(ori, nrm, k)
    ko = L._get_kappa_ori(cu(ori), cu(nrm), k)
    assert rel_err(ko.cpu().numpy(), kap_ori_o) < TOL
    ko32 = ko.cpu().numpy()
    fwd = O.geo_forward(adv, ori, nrm, ko32, k)

    a = cu(adv).requires_grad_(True)
    scale = torch.linspace(0.5, 2.0, b, device="cuda")
    total, cd, hd, cv = L.geo_loss(a, cu(ori), cu(nrm), ko, k, 1.0, 0.1, 1.0)
    for name, got in (("cd", cd), ("hd", hd), ("curv", cv)):
        assert rel_err(got.cpu().numpy(), fwd[name]) < TOL, name
    (total * scale).sum().backward()
    s = scale.cpu().numpy().astype(np.float64)
    G = O.geo_backward(adv, ori, fwd, ko32, s * 1.0, s * 0.1, s * 1.0)
    assert rel_err(a.grad.cpu().numpy(), G) < TOL
    # per-cloud check too (a big cloud must not hide a bad small one)
    for c in range(b):
        assert rel_err(a.grad[c].cpu().numpy(), G[c]) < 5 * TOL


def test_backward_is_deterministic():
    from geoa3_b200 import loss_utils as L

    adv, ori, nrm = make(6, 1024, 9, 2e-2)
    ko = L._get_kappa_ori(cu(ori), cu(nrm), 16)
    grads = []
    for _ in range(3):
        L.clear_cache()
        a = cu(adv).requires_grad_(True)
        L.geo_loss(a, cu(ori), cu(nrm), ko, 16)[0].sum().backward()
        grads.append(a.grad.clone())
    assert torch.equal(grads[0], grads[1]) and torch.equal(grads[0], grads[2])


@pytest.mark.parametrize("path", golden_files())
def test_reference_api_vs_golden(path):
    """The reference-named functions composed exactly as Attacker/geoA3_attack.py:131-162 does."""
    from geoa3_b200 import loss_utils as L

    g = np.load(path)
    k = int(g["k"])
    L.clear_cache()
    adv = cu(g["adv"]).requires_grad_(True)
    ori, nrm = cu(g["ori"]), cu(g["normal"])
    kap_ori = L._get_kappa_ori(ori, nrm, k)
    cd = L.chamfer_loss(adv, ori)
    hd = L.hausdorff_loss(adv, ori)
    kap_adv, nrm_adv = L._get_kappa_adv(adv, ori, nrm, k)
    cu_ = L.curvature_loss(adv, ori, kap_adv, kap_ori)
    (1.0 * cd + 0.1 * hd + 1.0 * cu_).sum().backward()
    assert rel_err(kap_ori.detach().cpu().numpy(), g["f64_kappa_ori"]) < TOL
    assert rel_err(kap_adv.detach().cpu().numpy(), g["f64_kappa_adv"]) < TOL
    assert np.array_equal(nrm_adv.cpu().numpy(), g["f32_nrm_adv"])
    for name, got in (("cd", cd), ("hd", hd), ("curv", cu_)):
        assert rel_err(got.detach().cpu().numpy(), g["f64_" + name]) < TOL, name
    assert rel_err(adv.grad.cpu().numpy(), g["f64_grad"]) < TOL
    # fused path gives the same numbers
    L.clear_cache()
    adv2 = cu(g["adv"]).requires_grad_(True)
    tot, cd2, hd2, cv2 = L.geo_loss(adv2, ori, nrm, kap_ori.detach(), k, 1.0, 0.1, 1.0)
    tot.sum().backward()
    assert rel_err(adv2.grad.cpu().numpy(), g["f64_grad"]) < TOL
    assert rel_err(cd2.cpu().numpy(), g["f64_cd"]) < TOL
    # one-sided chamfer (pseudo_chamfer_loss) value + grad vs oracle
    L.clear_cache()
    adv3 = cu(g["adv"]).requires_grad_(True)
    pc = L.pseudo_chamfer_loss(adv3, ori)
    pc.sum().backward()
    d1, j1 = O.nn1(g["adv"], g["ori"])
    assert rel_err(pc.detach().cpu().numpy(), d1.astype(np.float64).mean(1)) < TOL
    n = g["adv"].shape[2]
    Gp = 2.0 / n * (g["adv"].astype(np.float64) - np.take_along_axis(g["ori"].astype(np.float64), j1[:, None, :].repeat(3, 1), 2))
    assert rel_err(adv3.grad.cpu().numpy(), Gp) < TOL


def test_full_size_properties():
    """BASELINE config 2 size (B=250, N=1024): size-independent properties instead of an oracle sweep."""
    from geoa3_b200 import loss_utils as L
    from geoa3_b200 import ops

    b, n = 250, 1024
    pc, nr, _ = synth.make_batch(10, n, 0)
    reps = b // 10
    ori = np.tile(pc, (reps, 1, 1))
    nrm = np.tile(nr, (reps, 1, 1))
    adv = ori + synth.make_offsets(b, n, seed=0, std=1e-3)
    A, Oc = cu(adv), cu(ori)
    d1, j1, d2, i2 = ops.nn_pair(A, Oc)
    # tiny perturbation: every adv point's nearest original is its own source point (and vice versa)
    ar = torch.arange(n, device="cuda", dtype=torch.int32)[None].expand(b, n)
    assert (j1 == ar).float().mean() > 0.99 and (i2 == ar).float().mean() > 0.99
    # d1 is really the distance to the reported argmin
    gath = torch.gather(Oc, 2, j1.long()[:, None, :].expand(b, 3, n))
    assert torch.allclose(((A - gath) ** 2).sum(1), d1, rtol=1e-5, atol=1e-12)
    # self-kNN: first column is the point itself, distances ascending
    idx, dist = ops.knn(A, A, 17, return_dist=True)
    assert torch.equal(idx[:, :, 0], ar) and bool((dist[:, :, 1:] >= dist[:, :, :-1]).all())
    # spot-check 3 clouds against the oracle
    sel = [0, 123, 249]
    oi, _ = O.knn(adv[sel], adv[sel], 17)
    assert np.array_equal(idx[sel].cpu().numpy(), oi)
    # identical clouds (replicas) give identical losses/gradients: batch independence
    ko = L._get_kappa_ori(Oc, cu(nrm), 16)
    a = A.clone().requires_grad_(True)
    tot, cd, hd, cv = L.geo_loss(a, Oc, cu(nrm), ko, 16)
    tot.sum().backward()
    assert torch.equal(ko[:10], ko[10:20])
    a2 = A[:20].clone().requires_grad_(True)
    L.geo_loss(a2, Oc[:20].contiguous(), cu(nrm[:20]), ko[:20].contiguous(), 16)[0].sum().backward()
    assert torch.equal(a.grad[:20], a2.grad)
    # zero perturbation: CD = HD = 0 and the Chamfer/Hausdorff gradient vanishes
    z = Oc.clone().requires_grad_(True)
    (L.chamfer_loss(z, Oc) + L.hausdorff_loss(z, Oc)).sum().backward()
    assert float(z.grad.abs().max()) == 0.0


def test_hints_never_change_results():
    """Seeds / hinted thresholds are accelerators only: good, stale, adversarial and garbage hints all give
    the oracle's indices (including the in-place aliased form the attack driver uses)."""
    from geoa3_b200 import ops

    adv, ori, _ = make(3, 1000, 4, 5e-2)
    A, Oc = cu(adv), cu(ori)
    od1, oj1 = O.nn1(adv, ori)
    od2, oi2 = O.nn1(ori, adv)
    oi, od = O.knn(adv, adv, 17)
    rng = np.random.default_rng(0)
    garbage = torch.from_numpy(rng.integers(-5, 2000, (3, 1000)).astype(np.int32)).cuda()
    good_j, good_i = cu(oj1), cu(oi2)
    worst = torch.from_numpy(np.argmax(O.pairdist(adv, ori), 2).astype(np.int32)).cuda()
    for hj, hi in ((good_j, good_i), (garbage, garbage), (worst, worst), (None, garbage)):
        d1, j1, d2, i2 = ops.nn_pair(A, Oc, hint_a2o=hj, hint_o2a=hi)
        assert np.array_equal(j1.cpu().numpy(), oj1) and np.array_equal(i2.cpu().numpy(), oi2)
        assert np.array_equal(d1.cpu().numpy(), od1) and np.array_equal(d2.cpu().numpy(), od2)
    # in place: buffers are hint and output at once
    bj, bi_ = garbage.clone(), worst.clone()
    bd1, bd2 = torch.empty(3, 1000, device="cuda"), torch.empty(3, 1000, device="cuda")
    ops.nn_pair(A, Oc, hint_a2o=bj, hint_o2a=bi_, out=(bd1, bj, bd2, bi_))
    assert np.array_equal(bj.cpu().numpy(), oj1) and np.array_equal(bi_.cpu().numpy(), oi2)
    # kNN: exact hint, hint from a different (shifted) cloud, duplicated / garbage hint rows, far-away hint
    exact = cu(oi[:, :, 1:])
    other = cu(O.knn(ori, ori, 17)[0][:, :, 1:])
    dup = torch.zeros(3, 1000, 16, dtype=torch.int32, device="cuda")
    junk = torch.from_numpy(rng.integers(-3, 1500, (3, 1000, 16)).astype(np.int32)).cuda()
    far = torch.from_numpy(np.argsort(-O.pairdist(adv, adv), 2)[:, :, :16].astype(np.int32)).cuda()
    for h in (exact, other, dup, junk, far):
        idx, dist = ops.knn(A, A, 17, drop=0, return_dist=True, hint=h)
        assert np.array_equal(idx.cpu().numpy(), oi) and np.array_equal(dist.cpu().numpy(), od)
    buf = other.clone()
    ops.knn(A, A, 17, drop=1, hint=buf, out=buf)
    assert np.array_equal(buf.cpu().numpy(), oi[:, :, 1:])
    # lattice ties with hints
    pc, _ = synth.lattice_cloud(343)
    pc = np.stack([pc, pc * 0.5])
    oi2_, _ = O.knn(pc, pc, 9)
    h = cu(oi2_[:, :, 1:][:, :, ::-1].copy())
    idx, _ = ops.knn(cu(pc), cu(pc), 9, hint=h)
    assert np.array_equal(idx.cpu().numpy(), oi2_)


def test_attack_state_hints_match_unhinted():
    """The persistent hint buffers of the attack are accelerators only.  (i) WHAT they hold never matters: buffers
    carried over from the previous step, buffers left behind by a different batch and freshly initialised buffers
    give bitwise identical losses and gradients.  (ii) Against the plain reference API (no buffers: neighbour lists
    sorted by distance instead of in visiting order) the 1-NN indices and the neighbour SETS are identical; the
    kappa sums run in a different order, so values agree to rounding (1e-6), not bitwise."""
    from geoa3_b200 import loss_utils as L
    from geoa3_b200 import ops

    adv, ori, nrm = make(4, 1024, 7, 1e-2)
    other = make(4, 1024, 11, 5e-2)[0]
    ko = L._get_kappa_ori(cu(ori), cu(nrm), 16)
    hb, hb_stale = L.HintBuffers(), L.HintBuffers()
    for step in range(3):
        a_np = adv + synth.make_offsets(4, 1024, seed=10 + step, std=2e-3)
        # hb_stale saw an unrelated cloud last: its indices are valid seeds, but useless ones
        L.geo_loss(cu(other), cu(ori), cu(nrm), ko, 16, 1.0, 0.1, 1.0, hints=hb_stale)
        res = []
        for hints in (hb, hb_stale, L.HintBuffers(), None):
            L.clear_cache()
            a = cu(a_np).requires_grad_(True)
            tot, cd, hd, cv = L.geo_loss(a, cu(ori), cu(nrm), ko, 16, 1.0, 0.1, 1.0, hints=hints)
            tot.sum().backward()
            res.append((tot.detach().clone(), a.grad.clone()))
        for r in res[1:3]:
            assert torch.equal(res[0][0], r[0]) and torch.equal(res[0][1], r[1])
        assert rel_err(res[0][0].cpu().numpy(), res[3][0].cpu().numpy()) < 1e-6
        assert rel_err(res[0][1].cpu().numpy(), res[3][1].cpu().numpy()) < 1e-6
        A = cu(a_np)
        d1, j1, d2, i2 = ops.nn_pair(A, cu(ori))
        assert torch.equal(hb.jstar, j1) and torch.equal(hb.istar, i2) and torch.equal(hb_stale.jstar, j1)
        want = ops.knn(A, A, 17, drop=1)[0].sort(-1)[0]
        assert torch.equal(hb.nbr[16].sort(-1)[0], want) and torch.equal(hb_stale.nbr[16].sort(-1)[0], want)
        assert torch.equal(hb.nbr[16], hb_stale.nbr[16])   # same order, whatever the hint was


@pytest.mark.parametrize("scale,shift", [(1.0, 0.0), (1e-3, 0.0), (1e3, 0.0), (1.0, 50.0), (1e-2, 7.0), (30.0, -200.0)])
def test_nn_pair_filter_is_conservative(scale, shift):
    """The hot loop only FILTERS with |c|^2 - 2q.c; every magnitude / offset must still give the oracle's
    bit-exact (min, argmin): uncentred clouds make the expansion cancel catastrophically, near-duplicate
    points make the margin matter."""
    from geoa3_b200 import ops

    rng = np.random.default_rng(11)
    b, n, m = 3, 700, 900
    base = rng.standard_normal((b, 3, m)).astype(np.float32)
    ori = (base * scale + shift).astype(np.float32)
    pick = rng.integers(0, m, (b, n))
    adv = np.take_along_axis(ori, pick[:, None, :].repeat(3, 1), 2)
    adv = (adv + rng.standard_normal((b, 3, n)).astype(np.float32) * np.float32(scale * 1e-6)).astype(np.float32)
    adv[:, :, ::7] = np.take_along_axis(ori, pick[:, None, ::7].repeat(3, 1), 2)  # exact duplicates: ties at d = 0
    od1, oj1 = O.nn1(adv, ori)
    od2, oi2 = O.nn1(ori, adv)
    for hint in (None, torch.zeros(b, n, dtype=torch.int32, device="cuda")):
        d1, j1, d2, i2 = ops.nn_pair(cu(adv), cu(ori), hint_a2o=hint)
        assert np.array_equal(j1.cpu().numpy(), oj1) and np.array_equal(i2.cpu().numpy(), oi2)
        assert np.array_equal(d1.cpu().numpy(), od1) and np.array_equal(d2.cpu().numpy(), od2)


def test_nn_pair_any_visiting_order():
    """perm only changes the order in which points are visited (and enables bounding-box pruning when it is
    spatially coherent): identity, Morton and random permutations, with lattice ties, clustered duplicates and
    ragged sizes, must all return the oracle's result in ORIGINAL numbering."""
    from geoa3_b200 import ops

    rng = np.random.default_rng(21)
    cases = []
    adv, ori, _ = make(3, 1024, 4, 2e-2)
    cases.append((adv, ori))
    a2, _, _ = make(2, 777, 1, 1e-3)
    _, o2, _ = make(2, 2500, 6, 1e-3)
    cases.append((a2, o2))
    lat, _ = synth.lattice_cloud(343)
    cases.append((np.stack([lat, lat[:, ::-1].copy()]), np.stack([lat + 0.125, lat])))
    for adv, ori in cases:
        b, _, n = adv.shape
        m = ori.shape[2]
        od1, oj1 = O.nn1(adv, ori)
        od2, oi2 = O.nn1(ori, adv)
        A, Oc = cu(adv), cu(ori)
        orders = {"morton": (ops.morton_order(A)[0], ops.morton_order(Oc)[0]),
                  "random": (cu(np.stack([rng.permutation(n) for _ in range(b)]).astype(np.int32)),
                             cu(np.stack([rng.permutation(m) for _ in range(b)]).astype(np.int32)))}
        garbage = torch.from_numpy(rng.integers(0, min(n, m), (b, n)).astype(np.int32)).cuda()
        for name, (pa, po) in orders.items():
            inv = lambda p: torch.empty_like(p).scatter_(1, p.long(), torch.arange(p.shape[1], device="cuda", dtype=torch.int32).expand_as(p).contiguous())
            for kw in ({}, {"iperm_a": inv(pa), "iperm_o": inv(po), "hint_a2o": cu(oj1), "hint_o2a": cu(oi2)},
                       {"iperm_a": inv(pa), "iperm_o": inv(po), "hint_a2o": garbage}, {"hint_a2o": cu(oj1)}):
                d1, j1, d2, i2 = ops.nn_pair(A, Oc, perm_a=pa, perm_o=po, **kw)
                assert np.array_equal(j1.cpu().numpy(), oj1) and np.array_equal(i2.cpu().numpy(), oi2), (name, list(kw))
                assert np.array_equal(d1.cpu().numpy(), od1) and np.array_equal(d2.cpu().numpy(), od2), (name, list(kw))


def test_knn_any_visiting_order():
    """kNN with visiting orders (identity / Morton / random), hinted and unhinted, on smooth clouds, lattice ties
    and duplicated points: always the oracle's lexicographic (dist, original index) result."""
    from geoa3_b200 import ops

    rng = np.random.default_rng(31)
    clouds_ = [make(3, 1024, 2, 2e-2)[0], make(2, 1500, 5, 1e-2)[0]]
    lat, _ = synth.lattice_cloud(343)
    clouds_.append(np.stack([lat, lat * 0.5]))
    dup = clouds_[0][:2, :, :300].copy()
    dup[:, :, 150:200] = dup[:, :, 0:50]
    clouds_.append(np.ascontiguousarray(dup))
    for pts in clouds_:
        b, _, n = pts.shape
        P = cu(pts)
        pm, ipm = ops.morton_order(P)
        pr = cu(np.stack([rng.permutation(n) for _ in range(b)]).astype(np.int32))
        ipr = torch.empty_like(pr)
        ipr.scatter_(1, pr.long(), torch.arange(n, device="cuda", dtype=torch.int32).expand(b, n).contiguous())
        for K in (17, 33, 4):
            oi, od = O.knn(pts, pts, K)
            hint = cu(np.ascontiguousarray(oi[:, :, 1:]))
            for name, (pq, pc, ipc) in {"morton": (pm, pm, ipm), "random": (pr, pr, ipr), "noinv": (pm, pm, None),
                                       "cand_only": (None, pm, ipm)}.items():
                for h in (None, hint):
                    idx, dist = ops.knn(P, P, K, return_dist=True, hint=h, perm_q=pq, perm_c=pc, iperm_c=ipc)
                    assert np.array_equal(idx.cpu().numpy(), oi), (name, K, h is not None)
                    assert np.array_equal(dist.cpu().numpy(), od), (name, K, h is not None)


def test_api_ragged_sizes_and_small_k():
    """Reference API with n != m (Chamfer / Hausdorff between clouds of different sizes) and the reference's
    default k=2 (K=3 list), values and gradients against the oracle."""
    from geoa3_b200 import loss_utils as L

    rng = np.random.default_rng(8)
    adv = rng.standard_normal((2, 3, 300)).astype(np.float32) * 0.5
    ori = rng.standard_normal((2, 3, 517)).astype(np.float32) * 0.5
    a = cu(adv).requires_grad_(True)
    cd = L.chamfer_loss(a, cu(ori))
    hd = L.hausdorff_loss(a, cu(ori))
    (cd + 0.5 * hd).sum().backward()
    d1, j1 = O.nn1(adv, ori)
    d2, i2 = O.nn1(ori, adv)
    assert rel_err(cd.detach().cpu().numpy(), d1.astype(np.float64).mean(1) + d2.astype(np.float64).mean(1)) < TOL
    assert rel_err(hd.detach().cpu().numpy(), d1.max(1)) < TOL
    G = np.zeros((2, 3, 300))
    for c in range(2):
        A, Oo = adv[c].astype(np.float64), ori[c].astype(np.float64)
        G[c] += 2.0 / 300 * (A - Oo[:, j1[c]])
        for j in range(517):
            G[c][:, i2[c, j]] += 2.0 / 517 * (A[:, i2[c, j]] - Oo[:, j])
        ih = int(np.argmax(d1[c]))
        G[c][:, ih] += 0.5 * 2.0 * (A[:, ih] - Oo[:, j1[c, ih]])
    assert rel_err(a.grad.cpu().numpy(), G) < TOL
    # default k=2 curvature path
    pc, nr, _ = synth.make_batch(2, 200, 3)
    adv2 = pc + synth.make_offsets(2, 200, std=1e-2)
    ko = L._get_kappa_ori(cu(pc), cu(nr))
    kap_o, _ = O.kappa_ori(pc, nr, 2)
    assert rel_err(ko.cpu().numpy(), kap_o) < TOL
    a2 = cu(adv2).requires_grad_(True)
    ka, nrm_adv = L._get_kappa_adv(a2, cu(pc), cu(nr))
    cu_ = L.curvature_loss(a2, cu(pc), ka, ko)
    cu_.sum().backward()
    fwd = O.geo_forward(adv2, pc, nr, ko.cpu().numpy(), 2)
    assert rel_err(cu_.detach().cpu().numpy(), fwd["curv"]) < TOL
    Gc = O.geo_backward(adv2, pc, fwd, ko.cpu().numpy(), 0.0, 0.0, 1.0)
    assert rel_err(a2.grad.cpu().numpy(), Gc) < TOL


def _aux_calls(L, ori, nrm, k):
    return dict(displacement=lambda a: L.displacement_loss(a, ori, k),
                corr_normal=lambda a: L.corresponding_normal_loss(a, nrm, k),
                repulsion=lambda a: L.repulsion_loss(a, k, 0.03),
                kmean=lambda a: L.distance_kmean_loss(a, k),
                smoothing=lambda a: L.kNN_smoothing_loss(a, k),
                uniform=lambda a: L.uniform_loss(a))


@pytest.mark.parametrize("path", golden_files("aux_"), ids=lambda p: p.split("/")[-1])
def test_regularisers_vs_reference_golden(path):
    """Lib/loss_utils.py:99-190 on the CUDA neighbour searches vs the reference module's own outputs and
    autograd gradients (fixtures from tests/golden/make_golden.py).  Values 1e-5 relative; gradients are held
    to the accuracy the reference's own fp32 run achieves against its fp64 run (x4, floor 1e-5)."""
    from geoa3_b200 import loss_utils as L

    g = np.load(path)
    k = int(g["k"])
    ori, nrm = cu(g["ori"]), cu(g["normal"])
    for name, fn in _aux_calls(L, ori, nrm, k).items():
        a = cu(g["adv"]).requires_grad_(True)
        v = fn(a)
        v.sum().backward()
        assert tuple(v.shape) == tuple(g["f32_" + name].shape), name
        assert rel_err(v.detach().cpu().numpy(), g["f64_" + name]) < TOL, name
        ref64 = g["f64_" + name + "_grad"]
        bar = max(TOL, 4 * rel_err(g["f32_" + name + "_grad"], ref64))
        assert rel_err(a.grad.cpu().numpy(), ref64) < bar, (name, bar)


def test_regularisers_vs_oracle_seeded():
    """Same functions at other sizes / k against the numpy restatement (values), incl. a K=33 search."""
    from geoa3_b200 import loss_utils as L

    for (b, n, k, std) in ((3, 1024, 16, 1e-2), (2, 700, 32, 5e-2), (4, 200, 2, 1e-3)):
        adv, ori, nrm = make(b, n, 2, std)
        want = O.aux_losses(adv, ori, nrm, k)
        calls = _aux_calls(L, cu(ori), cu(nrm), k)
        for name, w in want.items():
            assert rel_err(calls[name](cu(adv)).cpu().numpy(), w) < TOL, (name, n, k)
        assert abs(float(L.uniform_loss(cu(adv))) - O.uniform_loss(adv)) < TOL * O.uniform_loss(adv)


def test_defense_filters_vs_reference_golden():
    """defense.py:25-40 (statistical outlier removal) on the CUDA top-K search vs the reference's own function
    outputs: the same points are kept (bit-identical kept clouds), the same counts are returned."""
    import os.path as osp

    from geoa3_b200 import defense as D
    from helpers import GOLDEN_DIR

    g = np.load(osp.join(GOLDEN_DIR, "defense_cases.npz"))
    i = 0
    while "c%d_pc" % i in g:
        pc = cu(g["c%d_pc" % i])
        drop, alpha, knn = g["c%d_args" % i]
        out, num = D.point_removal_fn(pc, "outliers_variance", int(drop), float(alpha), int(knn))
        assert num == int(g["c%d_var_num" % i]) and np.array_equal(out.cpu().numpy(), g["c%d_var_pc" % i])
        out, num = D.point_removal_fn(pc, "outliers_fixNum", int(drop), float(alpha), int(knn))
        assert num == int(g["c%d_fix_num" % i]) and np.array_equal(out.cpu().numpy(), g["c%d_fix_pc" % i])
        out, num = D.point_removal_fn(pc, "rand_drop", int(drop), float(alpha), int(knn))
        assert num == int(drop) and out.shape == (1, 3, pc.shape[2] - int(drop))
        i += 1
    assert i == 3


def test_backward_large_clouds():
    """Clouds whose neighbour lists outgrow shared memory (n*k > 65535 edges: BASELINE config[3] sizes) take the
    global-workspace backward: gradients vs the oracle, and bit-identical to the fused kernel where both apply."""
    import os

    from geoa3_b200 import loss_utils as L
    from geoa3_b200 import ops

    for (b, n, k, std) in ((2, 4096, 16, 1e-2), (1, 2500, 32, 3e-2)):
        adv, ori, nrm = make(b, n, 1, std)
        assert ops._lib.load().geoa3_loss_bwd_workspace_bytes(b, n, n, k) > 0
        a = cu(adv).requires_grad_(True)
        ko = L._get_kappa_ori(cu(ori), cu(nrm), k)
        tot, cd, hd, cur = L.geo_loss(a, cu(ori), cu(nrm), ko, k, 1.0, 0.1, 1.0)
        tot.sum().backward()
        ko_o, _ = O.kappa_ori(ori, nrm, k)
        fwd = O.geo_forward(adv, ori, nrm, ko_o, k)
        g = np.ones(b)
        want = O.geo_backward(adv, ori, fwd, ko_o, g * 1.0, g * 0.1, g * 1.0)
        assert rel_err(a.grad.cpu().numpy(), want) < TOL
        L.clear_cache()
    # same inputs through both paths
    adv, ori, nrm = make(3, 700, 4, 2e-2)
    grads = []
    for force in (False, True):
        if force:
            os.environ["GEOA3_BWD_LARGE"] = "1"
        try:
            a = cu(adv).requires_grad_(True)
            ko = L._get_kappa_ori(cu(ori), cu(nrm), 16)
            L.geo_loss(a, cu(ori), cu(nrm), ko, 16, 1.0, 0.1, 1.0)[0].sum().backward()
            grads.append(a.grad.clone())
        finally:
            os.environ.pop("GEOA3_BWD_LARGE", None)
        L.clear_cache()
    assert torch.equal(grads[0], grads[1])


def test_pca_estimators_vs_reference_golden():
    """estimate_normal / estimate_perpendicular / estimate_normal_via_ori_normal (Lib/utility.py:40-149) on the CUDA
    neighbour search.  Normals vs the reference function executed in place (fixture), compared up to sign — the
    reference's own orientation rule reads rounding noise — and allowing the few points whose two smallest
    eigenvalues nearly coincide; the jitter must lie in the tangent plane and respect the clip."""
    import os.path as osp

    from geoa3_b200 import utility as U
    from helpers import GOLDEN_DIR

    g = np.load(osp.join(GOLDEN_DIR, "estimate_normal_cases.npz"))
    i = 0
    while "c%d_pc" % i in g:
        pc, k, ref = g["c%d_pc" % i], int(g["c%d_k" % i]), g["c%d_normal" % i]
        got = U.estimate_normal(cu(pc), k).cpu().numpy()
        assert got.shape == ref.shape
        # the reference's sign rule is -sign(sum of centred neighbours . normal): rounding noise, and exactly 0 (a
        # zero "normal") wherever that sum cancels exactly — on either side, at different points
        live = (np.linalg.norm(got, axis=1) > 0.5) & (np.linalg.norm(ref, axis=1) > 0.5)
        assert live.mean() > 0.8
        cos = np.abs((got * ref).sum(1))[live]
        assert np.mean(cos > 0.999) > 0.97 and np.median(cos) > 0.99999, (np.mean(cos > 0.999), np.median(cos))
        torch.manual_seed(0)
        jit = U.estimate_perpendicular(cu(pc), k, sigma=0.01, clip=0.05).cpu().numpy()
        assert jit.shape == pc.shape and np.abs(jit).max() <= 0.1 + 1e-6
        along = (np.abs((jit * got).sum(1)) / (np.linalg.norm(jit, axis=1) + 1e-12))[live]
        assert np.median(along) < 1e-3                    # tangent-plane noise: no component along the normal
        i += 1
    assert i == 2
    pc, nr, _ = synth.make_batch(2, 400, 0)
    same = U.estimate_normal_via_ori_normal(cu(pc), cu(pc), cu(nr), 3)
    assert torch.equal(same, cu(nr))                       # unmoved points take the original normal itself
    moved = pc + synth.make_offsets(2, 400, std=2e-2)
    est = U.estimate_normal_via_ori_normal(cu(moved), cu(pc), cu(nr), 3).cpu().numpy()
    assert np.allclose(np.linalg.norm(est, axis=1), 1.0, atol=1e-4) and np.median(np.abs((est * nr).sum(1))) > 0.95


def test_knn_set_members_exact():
    """geoa3_knn_set (the neighbour lists of the curvature term): bit-exact MEMBERSHIP and member distances against
    the oracle's top-K (minus the dropped self match) for every kind of hint and visiting order, on smooth clouds,
    lattice ties, duplicated points, ragged sizes, multi-chunk clouds and clouds with fewer points than the list
    capacity; the order it writes (ascending visiting position) does not depend on the hint."""
    from geoa3_b200 import ops

    rng = np.random.default_rng(5)
    lat, _ = synth.lattice_cloud(343)
    dup = make(2, 300, 4, 1e-2)[0]
    dup[:, :, 150:200] = dup[:, :, 0:50]
    cases = [(make(3, 1024, 2, 2e-2)[0], 17), (make(2, 1024, 1, 1e-2)[0], 33), (make(2, 777, 5, 5e-2)[0], 9),
             (np.stack([lat, lat * 0.5]), 9), (np.ascontiguousarray(dup), 17), (make(1, 2500, 3, 1e-2)[0], 17),
             (make(2, 40, 0, 1e-1)[0], 33), (make(2, 20, 0, 1e-1)[0], 17)]
    for pts, K in cases:
        b, _, n = pts.shape
        K = min(K, n)
        P = cu(pts)
        oi, od = O.knn(pts, pts, K)
        want_i, want_d = oi[:, :, 1:], od[:, :, 1:]

        def check(idx, dist, tag):
            i_, d_ = idx.cpu().numpy(), dist.cpu().numpy()
            key = d_.view(np.int32).astype(np.int64) * 65536 + i_
            o = np.argsort(key, -1)
            assert np.array_equal(np.take_along_axis(i_, o, -1), want_i), (n, K, tag)
            assert np.array_equal(np.take_along_axis(d_, o, -1), want_d), (n, K, tag)

        exact = cu(want_i)
        stale = cu(O.knn(pts + 0.05, pts[:, :, ::-1].copy(), K)[0][:, :, 1:])
        junk = cu(rng.integers(-3, n + 50, (b, n, K - 1)).astype(np.int32))
        zeros = torch.zeros(b, n, K - 1, dtype=torch.int32, device="cuda")
        base = ops.knn(P, P, K, drop=1, return_dist=True, members_only=True)
        check(*base, "unhinted")
        pm, ipm = ops.visit_order(P)
        pr = cu(np.stack([rng.permutation(n) for _ in range(b)]).astype(np.int32))
        ipr = torch.empty_like(pr)
        ipr.scatter_(1, pr.long(), torch.arange(n, device="cuda", dtype=torch.int32).expand(b, n).contiguous())
        for h, tag in ((exact, "exact"), (stale, "stale"), (junk, "junk"), (zeros, "dup")):
            idx, dist = ops.knn(P, P, K, drop=1, return_dist=True, hint=h, members_only=True)
            check(idx, dist, tag)
            assert torch.equal(idx, base[0]), (n, K, tag, "order depends on the hint")
            for perm, iperm, ptag in ((pm, ipm, "visit"), (pr, ipr, "random")):
                i2, d2 = ops.knn(P, P, K, drop=1, return_dist=True, hint=h, perm_q=perm, perm_c=perm, iperm_c=iperm,
                                 members_only=True)
                check(i2, d2, tag + "+" + ptag)
        # in place: the buffer is hint and output at once, three refreshes of a moving cloud
        buf, cur = exact.clone(), pts
        for step in range(3):
            cur = (cur + synth.make_offsets(b, n, seed=40 + step, std=3e-3)).astype(np.float32)
            C = cu(cur)
            ops.knn(C, C, K, drop=1, hint=buf, out=buf, perm_q=pm, perm_c=pm, iperm_c=ipm, members_only=True)
            assert np.array_equal(np.sort(buf.cpu().numpy(), -1), np.sort(O.knn(cur, cur, K)[0][:, :, 1:], -1))
    # drop=0 and a cross query (query cloud != candidate cloud)
    q, c = make(2, 333, 2, 1e-1)[0], make(2, 777, 6, 1e-2)[0]
    oi, od = O.knn(q, c, 5)
    idx, dist = ops.knn(cu(q), cu(c), 5, drop=0, return_dist=True, members_only=True)
    assert np.array_equal(np.sort(idx.cpu().numpy(), -1), np.sort(oi, -1))


@pytest.mark.parametrize("path", golden_files())
def test_gradient_elementwise_tolerance(path):
    """north_star's 1e-5 RELATIVE on every gradient entry, not only norm-wise: |got - ref| <= 1e-5*|ref| for entries
    down to 1 % of the cloud's largest, with the absolute floor 1e-5 * 1e-2 * max|ref| below that (helpers.elem_err).
    Reference = the fp64 autograd gradient of the reference's own functions (golden fixtures)."""
    from helpers import elem_err

    from geoa3_b200 import loss_utils as L

    g = np.load(path)
    k = int(g["k"])
    L.clear_cache()
    adv = cu(g["adv"]).requires_grad_(True)
    ori, nrm = cu(g["ori"]), cu(g["normal"])
    ko = L._get_kappa_ori(ori, nrm, k)
    (L.chamfer_loss(adv, ori) + 0.1 * L.hausdorff_loss(adv, ori)
     + L.curvature_loss(adv, ori, L._get_kappa_adv(adv, ori, nrm, k)[0], ko)).sum().backward()
    got = adv.grad.cpu().numpy()
    for c in range(got.shape[0]):   # per cloud: a big cloud must not lend its scale to a small one
        assert elem_err(got[c], g["f64_grad"][c]) <= 1.0, (c, elem_err(got[c], g["f64_grad"][c]))
    # the fused node delivers the same gradient under the same bound
    L.clear_cache()
    a2 = cu(g["adv"]).requires_grad_(True)
    L.geo_loss(a2, ori, nrm, ko, k, 1.0, 0.1, 1.0, hints=L.HintBuffers())[0].sum().backward()
    for c in range(got.shape[0]):
        assert elem_err(a2.grad[c].cpu().numpy(), g["f64_grad"][c]) <= 1.0
    for name, val in (("cd", L.chamfer_loss(adv, ori)), ("hd", L.hausdorff_loss(adv, ori))):
        assert elem_err(val.detach().cpu().numpy(), g["f64_" + name], floor=0.0) <= 1.0, name
    assert elem_err(L._get_kappa_adv(adv, ori, nrm, k)[0].detach().cpu().numpy(), g["f64_kappa_adv"]) <= 1.0


def test_full_batch_index_parity_B250():
    """BASELINE config[1]/[2] size, EVERY cloud against the C oracle (not a spot check): fused 1-NN both directions,
    kNN K=17 (sorted kernel and member-set kernel), FPS 1024->512 and ball_query r=0.2 / nsample=64."""
    from geoa3_b200 import ops

    b, n = 250, 1024
    pc, _, _ = synth.make_batch(50, n, 0)
    ori = np.tile(pc, (5, 1, 1))
    adv = (ori + synth.make_offsets(b, n, seed=3, std=1e-2)).astype(np.float32)
    A, Oc = cu(adv), cu(ori)
    d1, j1, d2, i2 = ops.nn_pair(A, Oc)
    od1, oj1 = O.nn1(adv, ori)
    od2, oi2 = O.nn1(ori, adv)
    assert np.array_equal(j1.cpu().numpy(), oj1) and np.array_equal(i2.cpu().numpy(), oi2)
    assert np.array_equal(d1.cpu().numpy(), od1) and np.array_equal(d2.cpu().numpy(), od2)
    pm, ipm = ops.visit_order(Oc)
    d1p, j1p, d2p, i2p = ops.nn_pair(A, Oc, hint_a2o=j1, hint_o2a=i2, perm_a=pm, perm_o=pm, iperm_a=ipm, iperm_o=ipm)
    assert torch.equal(j1p, j1) and torch.equal(i2p, i2) and torch.equal(d1p, d1)
    oi, od = O.knn(adv, adv, 17)
    idx, dist = ops.knn(A, A, 17, return_dist=True)
    assert np.array_equal(idx.cpu().numpy(), oi) and np.array_equal(dist.cpu().numpy(), od)
    hint = cu(O.knn(ori, ori, 17)[0][:, :, 1:])   # "previous step" = the unperturbed cloud
    mem = ops.knn(A, A, 17, drop=1, hint=hint, perm_q=pm, perm_c=pm, iperm_c=ipm, members_only=True)[0]
    assert np.array_equal(np.sort(mem.cpu().numpy(), -1), np.sort(oi[:, :, 1:], -1))
    xyz = np.ascontiguousarray(adv.transpose(0, 2, 1))
    X = cu(xyz)
    fi = ops.furthest_point_sampling(X, 512)
    ofi = O.fps(xyz, 512)
    assert np.array_equal(fi.cpu().numpy(), ofi)
    new = np.ascontiguousarray(np.take_along_axis(xyz, ofi[:, :, None].astype(np.int64).repeat(3, 2), 1))
    bq = ops.ball_query(cu(new), X, 0.2, 64)
    assert np.array_equal(bq.cpu().numpy(), O.ball_query(new, xyz, 0.2, 64))


def test_graph_replay_final_buffers_match_oracle():
    """60 replays of the captured attack step (persistent hint buffers refreshed in place, slab-ordered pruned
    searches, the product configuration): the index buffers left behind by the LAST replay are the oracle's answer for
    the cloud that replay saw — a stale-hint or replay-aliasing bug would need many steps to show and would show here."""
    from geoa3_b200 import attack as atk
    from geoa3_b200.victims import build_victim

    torch.manual_seed(0)
    net = build_victim("PointNet").cuda().eval()
    for p in net.parameters():
        p.requires_grad_(False)
    b, n, k = 6, 1024, 16
    pc, nr, lab = synth.make_batch(b, n, 4)
    dev = torch.device("cuda")
    cfg = atk.make_cfg(binary_max_steps=1, iter_max_steps=100, curv_loss_knn=k)
    lab_t = torch.from_numpy(lab).to(dev)
    st = atk.AttackState(net, torch.from_numpy(pc).to(dev), torch.from_numpy(nr).to(dev), lab_t, lab_t, cfg, targeted=False)
    init = atk.default_offsets(b, n, 0, 0).to(dev)
    st.begin_search_step(0, init)
    st.capture()
    st.reset_global()
    st.begin_search_step(0, init)
    for _ in range(60):
        st.run_step()
    torch.cuda.synchronize()
    adv = st.last["adv"].cpu().numpy()          # the cloud the last replay searched (before its Adam update)
    assert np.abs(adv - pc).max() > 5e-3         # the optimisation really moved the cloud
    od1, oj1 = O.nn1(adv, pc)
    od2, oi2 = O.nn1(pc, adv)
    hb = st.hints
    assert np.array_equal(hb.jstar.cpu().numpy(), oj1) and np.array_equal(hb.istar.cpu().numpy(), oi2)
    assert np.array_equal(hb.d1.cpu().numpy(), od1) and np.array_equal(hb.d2.cpu().numpy(), od2)
    oi, _ = O.knn(adv, adv, k + 1)
    assert np.array_equal(np.sort(hb.nbr[k].cpu().numpy(), -1), np.sort(oi[:, :, 1:], -1))
    # and the loss values of that replay agree with the oracle on that cloud
    ko, _ = O.kappa_ori(pc, nr, k)
    fwd = O.geo_forward(adv, pc, nr, ko, k)
    assert rel_err(st.last["dis"].cpu().numpy(), fwd["cd"]) < TOL and rel_err(st.last["hd"].cpu().numpy(), fwd["hd"]) < TOL
    assert rel_err(st.last["curv"].cpu().numpy(), fwd["curv"]) < TOL


@pytest.mark.parametrize("n,k", [(2970, 16), (3000, 16), (3740, 8), (3800, 8), (2690, 20), (2720, 20), (1024, 32)])
def test_backward_fused_kernel_near_shared_memory_limit(n, k):
    """Cloud sizes around the point where the fused backward's shared-memory plan stops fitting (its counter array used
    to be sized for 16-bit counters although the CSR builder uses 32-bit ones — harmless at n = 1024, corrupting in
    n = 2977..3055 at k = 16, 3745..3869 at k = 8, 2700..2763 at k = 20): the largest size that still takes the fused
    kernel and a size inside each formerly corrupting range; gradient vs the oracle, bit-identical to the
    global-workspace path."""
    import os

    from geoa3_b200 import loss_utils as L
    from geoa3_b200 import ops

    b = 2
    adv, ori, nrm = make(b, n, 2, 2e-2)
    fused = ops._lib.load().geoa3_loss_bwd_workspace_bytes(b, n, n, k) == 0
    assert fused == ((n, k) in ((2970, 16), (3740, 8), (2690, 20), (1024, 32)))
    ko = L._get_kappa_ori(cu(ori), cu(nrm), k)
    grads = []
    for env in (None, "1"):
        if env:
            os.environ["GEOA3_BWD_LARGE"] = env
        try:
            L.clear_cache()
            a = cu(adv).requires_grad_(True)
            L.geo_loss(a, cu(ori), cu(nrm), ko, k, 1.0, 0.1, 1.0)[0].sum().backward()
            grads.append(a.grad.clone())
        finally:
            os.environ.pop("GEOA3_BWD_LARGE", None)
    assert torch.equal(grads[0], grads[1])
    ko_o, _ = O.kappa_ori(ori, nrm, k)
    fwd = O.geo_forward(adv, ori, nrm, ko_o, k)
    g = np.ones(b)
    want = O.geo_backward(adv, ori, fwd, ko_o, g * 1.0, g * 0.1, g * 1.0)
    assert rel_err(grads[0].cpu().numpy(), want) < TOL
