"""GPU parity (-m gpu) of the cell-grid searches — geoa3_cell_sort + geoa3_knn_cells / geoa3_nn_pair_cells — through the
C ABI against the C oracle: bit-exact members / distances / argmins for every kind of hint, several grid sizes and the
awkward clouds (lattice ties, duplicates, far-shifted, tiny, flat, single-valued, ragged, n != m, disjoint boxes), the
blob layout itself, and the product path (HintBuffers with and without cells give the same losses and indices)."""
import numpy as np
import pytest
import torch

from oracle import oracle as O
from oracle import synth

pytestmark = pytest.mark.gpu


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def make(b, n, seed, std):
    pc, nr, _ = synth.make_batch(b, n, seed)
    return (pc + synth.make_offsets(b, n, seed=seed + 100, std=std)).astype(np.float32), pc, nr


def knn_cases():
    lat, _ = synth.lattice_cloud(343)
    dup = make(2, 300, 4, 1e-2)[0]
    dup[:, :, 150:200] = dup[:, :, 0:50]
    flat = make(2, 400, 9, 1e-2)[0]
    flat[:, 2, :] = 0.25
    return [("smooth17", make(3, 1024, 2, 2e-2)[0], 17), ("smooth33", make(2, 1024, 1, 1e-2)[0], 33),
            ("ragged9", make(2, 777, 5, 5e-2)[0], 9), ("lattice", np.stack([lat, lat * 0.5]), 9),
            ("duplicates", np.ascontiguousarray(dup), 17), ("n2500", make(1, 2500, 3, 1e-2)[0], 17),
            ("n40_K33", make(2, 40, 0, 1e-1)[0], 33), ("n20", make(2, 20, 0, 1e-1)[0], 17),
            ("shifted", make(2, 500, 7, 1e-2)[0] + np.float32(100.0), 17),
            ("tiny", make(2, 500, 8, 1e-2)[0] * np.float32(1e-3), 17), ("flat", flat, 17),
            ("single_value", np.zeros((1, 3, 64), np.float32) + np.float32(0.5), 9),
            ("K11_between_classes", make(2, 1024, 13, 1e-2)[0], 11), ("n4096_unstaged", make(1, 4096, 11, 1e-2)[0], 17)]


@pytest.mark.parametrize("name,pts,K", knn_cases(), ids=[c[0] for c in knn_cases()])
def test_knn_cells_members_exact(name, pts, K):
    from geoa3_b200 import ops

    rng = np.random.default_rng(5)
    b, _, n = pts.shape
    K = min(K, n)
    P = cu(pts)
    oi, od = O.knn(pts, pts, K)
    want_i, want_d = oi[:, :, 1:], od[:, :, 1:]
    exact = cu(want_i)
    stale = cu(O.knn(pts + 0.05, pts[:, :, ::-1].copy(), K)[0][:, :, 1:])
    junk = cu(rng.integers(-3, n + 50, (b, n, K - 1)).astype(np.int32))
    zeros = torch.zeros(b, n, K - 1, dtype=torch.int32, device="cuda")
    for G in [1, 2, (5, 1, 3), None, 13, (9, 1, 12), ("kref", 3.0), ("kref", 60.0)]:
        cells = ops.cell_sort(P, kref=G[1]) if isinstance(G, tuple) and G[0] == "kref" else ops.cell_sort(P, kref=K, grid=G)
        first = None
        for h, tag in ((None, "none"), (exact, "exact"), (stale, "stale"), (junk, "junk"), (zeros, "dup")):
            idx, dist = ops.knn_cells(cells, K, drop=1, return_dist=True, hint=h)
            i_, d_ = idx.cpu().numpy(), dist.cpu().numpy()
            o = np.argsort(d_.view(np.int32).astype(np.int64) * 65536 + i_, -1)
            assert np.array_equal(np.take_along_axis(i_, o, -1), want_i), (name, G, tag)
            assert np.array_equal(np.take_along_axis(d_, o, -1), want_d), (name, G, tag)
            first = idx if first is None else first
            assert torch.equal(idx, first), (name, G, tag, "member order depends on the hint")
    # drop = 0 keeps the self match; in place: the buffer is hint and output at once, three refreshes of a moving cloud
    idx0 = ops.knn_cells(ops.cell_sort(P, kref=K), K, drop=0)[0]
    assert np.array_equal(np.sort(idx0.cpu().numpy(), -1), np.sort(oi, -1))
    buf, cur = exact.clone(), pts
    for step in range(3):
        cur = (cur + synth.make_offsets(b, n, seed=40 + step, std=3e-3)).astype(np.float32)
        ops.knn_cells(ops.cell_sort(cu(cur), kref=K), K, drop=1, hint=buf, out=buf)
        assert np.array_equal(np.sort(buf.cpu().numpy(), -1), np.sort(O.knn(cur, cur, K)[0][:, :, 1:], -1))


def nn_cases():
    lat, _ = synth.lattice_cloud(343)
    d = list(make(2, 300, 4, 1e-2)[:2])
    d[1] = d[1].copy()
    d[1][:, :, 150:200] = d[1][:, :, 0:50]
    return [("smooth", *make(3, 1024, 2, 2e-2)[:2]), ("ragged", *make(2, 1000, 1, 5e-2)[:2]),
            ("n_ne_m", make(2, 333, 2, 1e-1)[0], make(2, 777, 6, 1e-2)[1]),
            ("lattice_ties", np.stack([lat, lat * 0.5]), np.stack([lat * 0.5, lat])),
            ("shifted", *[x + np.float32(100.0) for x in make(2, 500, 7, 1e-2)[:2]]),
            ("tiny", *[x * np.float32(1e-3) for x in make(2, 500, 8, 1e-2)[:2]]),
            ("disjoint_boxes", make(2, 400, 9, 1e-2)[0] + np.float32(3.0), make(2, 400, 9, 1e-2)[1]),
            ("single_value_queries", np.zeros((1, 3, 64), np.float32) + np.float32(0.5), make(1, 90, 3, 1e-2)[1]),
            ("duplicated_candidates", d[0], d[1]), ("n20", *make(2, 20, 0, 1e-1)[:2]), ("n4096", *make(1, 4096, 11, 1e-2)[:2])]


@pytest.mark.parametrize("name,adv,ori", nn_cases(), ids=[c[0] for c in nn_cases()])
def test_nn_pair_cells_bitexact(name, adv, ori):
    from geoa3_b200 import ops

    rng = np.random.default_rng(7)
    adv, ori = np.ascontiguousarray(adv, np.float32), np.ascontiguousarray(ori, np.float32)
    b, _, n = adv.shape
    m = ori.shape[2]
    od1, oj1 = O.nn1(adv, ori)
    od2, oi2 = O.nn1(ori, adv)
    A, Oc = cu(adv), cu(ori)
    for ga, go in ((1, 1), ((3, 1, 4), 5), (None, None), (13, (11, 1, 7))):
        ba, bo = ops.cell_sort(A, kref=17, grid=ga), ops.cell_sort(Oc, kref=4, grid=go)
        hints = [(None, None, "none"), (cu(oj1), cu(oi2), "exact"),
                 (cu(rng.integers(-3, m + 50, (b, n)).astype(np.int32)), cu(rng.integers(-3, n + 50, (b, m)).astype(np.int32)), "junk"),
                 (torch.zeros(b, n, dtype=torch.int32, device="cuda"), torch.zeros(b, m, dtype=torch.int32, device="cuda"), "zeros")]
        for h1, h2, tag in hints:
            d1, j1, d2, i2 = ops.nn_pair_cells(ba, bo, hint_a2o=h1, hint_o2a=h2)
            assert np.array_equal(j1.cpu().numpy(), oj1) and np.array_equal(i2.cpu().numpy(), oi2), (name, ga, go, tag)
            assert np.array_equal(d1.cpu().numpy(), od1) and np.array_equal(d2.cpu().numpy(), od2), (name, ga, go, tag)
        d1, j1, d2, i2 = ops.nn_pair_cells(ba, bo, both=False)
        assert d2 is None and np.array_equal(j1.cpu().numpy(), oj1) and np.array_equal(d1.cpu().numpy(), od1)
        # in place: hint and output are the same buffer
        j, i = cu(oj1).clone(), cu(oi2).clone()
        ops.nn_pair_cells(ba, bo, hint_a2o=j, hint_o2a=i, out=(d1, j, torch.empty_like(i, dtype=torch.float32), i))
        assert np.array_equal(j.cpu().numpy(), oj1) and np.array_equal(i.cpu().numpy(), oi2)


def test_cell_blob_layout():
    """The blob geoa3_cell_sort writes: a permutation of the cloud in cell-major order (ascending original index inside
    a cell), start table consistent with the cell of every stored point, inverse permutation, and bitwise run-to-run
    reproducibility (the sort hands out slots with shared atomics; the per-cell ordering pass makes it deterministic)."""
    from geoa3_b200 import ops

    pts = make(3, 1000, 2, 2e-2)[0]
    n = 1000
    P = cu(pts)
    for grid in ((7, 5, 6), None):
        _check_blob(ops, pts, P, n, grid)


def _check_blob(ops, pts, P, n, grid):
    cells = ops.cell_sort(P, kref=17, grid=grid)
    blobs, nc = cells.blobs, cells.ncap
    assert blobs.shape[1] == ops._lib.load().geoa3_cell_blob_bytes(n, nc) and blobs.shape[1] % 16 == 0
    for _ in range(3):
        assert torch.equal(ops.cell_sort(P, kref=17, grid=grid).blobs, blobs)
    again = ops.cell_sort(P, kref=17, grid=grid, out=cells)
    assert again is cells
    raw = blobs.cpu().numpy()
    for c in range(pts.shape[0]):
        gp = raw[c, :64].view(np.float32)
        G = tuple(int(g) for g in gp[13:16])
        assert np.array_equal(gp[10:13], np.float32(G) - 1) and G[0] * G[1] * G[2] <= nc
        if grid is not None:
            assert G == tuple(grid)
        ncell = G[0] * G[1] * G[2]
        cl = raw[c, 64:64 + 16 * n].view(np.float32).reshape(n, 4)
        orig = cl[:, 3].view(np.int32)
        cs0 = 64 + 16 * n
        cstart = raw[c, cs0:cs0 + 2 * (nc + 1)].view(np.uint16).astype(np.int64)
        ip0 = cs0 + 2 * ((nc + 1 + 7) // 8 * 8)
        ipos = raw[c, ip0:ip0 + 2 * n].view(np.uint16)
        assert np.array_equal(np.sort(orig), np.arange(n))
        assert np.array_equal(cl[:, :3], pts[c].T[orig])
        assert np.array_equal(ipos[orig], np.arange(n))
        assert cstart[0] == 0 and cstart[nc] == n and np.all(np.diff(cstart) >= 0) and np.all(cstart[ncell:] == n)
        lo, inv_h = gp[0:3], gp[3:6]
        cell = np.clip(((cl[:, :3] - lo) * inv_h).astype(np.float32), 0, np.float32(G) - 1).astype(np.int64)
        key = (cell[:, 2] * G[1] + cell[:, 1]) * G[0] + cell[:, 0]
        assert np.all(np.diff(key) >= 0)
        assert np.array_equal(np.searchsorted(key, np.arange(nc + 1), side="left"), cstart)
        same = np.diff(key) == 0
        assert np.all(np.diff(orig)[same] > 0)


def test_hint_buffers_cells_equal_scan():
    """The product path (fused loss through persistent HintBuffers) with the cell-grid searches and with the scanning
    kernels: identical index buffers (as sets per row for the neighbour lists), loss values within rounding of the
    summation order, over several steps of a moving cloud."""
    from geoa3_b200 import loss_utils as L

    b, n, k = 4, 1024, 16
    adv0, ori, nrm = make(b, n, 3, 1e-2)
    Oc, Nr = cu(ori), cu(nrm)
    ko = L._get_kappa_ori(Oc, Nr, k)
    hc, hs = L.HintBuffers(use_cells=True), L.HintBuffers(use_cells=False)
    cur = adv0
    for step in range(4):
        cur = (cur + synth.make_offsets(b, n, seed=70 + step, std=4e-3)).astype(np.float32)
        outs = []
        for hb in (hc, hs):
            L.clear_cache()
            a = cu(cur).requires_grad_(True)
            tot, cd, hd, cv = L.geo_loss(a, Oc, Nr, ko, k, 1.0, 0.1, 1.0, hints=hb)
            tot.sum().backward()
            outs.append((tot.detach(), cd, hd, cv, a.grad.clone()))
        assert hc.cells_adv is not None and hs.cells_adv is None
        assert torch.equal(hc.jstar, hs.jstar) and torch.equal(hc.istar, hs.istar)
        assert torch.equal(hc.d1, hs.d1) and torch.equal(hc.d2, hs.d2)
        assert torch.equal(hc.nbr[k].sort(-1)[0], hs.nbr[k].sort(-1)[0])
        oi = O.knn(cur, cur, k + 1)[0][:, :, 1:]
        assert np.array_equal(np.sort(hc.nbr[k].cpu().numpy(), -1), np.sort(oi, -1))
        assert torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][2], outs[1][2])   # CD, HD: same inputs
        for x, y in zip(outs[0], outs[1]):
            assert float((x - y).abs().max()) <= 2e-6 * float(y.abs().max()) + 1e-12


@pytest.mark.parametrize("b,n,k,std,single,w_cu", [(4, 1024, 16, 1e-2, False, 1.0), (3, 1000, 16, 3e-2, False, 1.0),
                                                   (2, 512, 32, 1e-2, True, 1.0), (2, 200, 2, 1e-1, False, 1.0),
                                                   (2, 2970, 16, 1e-2, False, 1.0), (3, 1024, 16, 1e-2, False, 0.0),
                                                   (2, 777, 16, 1e-2, True, 0.0)])
def test_fused_fwd_bwd_equals_two_kernels(b, n, k, std, single, w_cu):
    """geoa3_geo_fwd_bwd (kappa + reductions + unit-upstream gradient in one launch, scaled in backward) against the
    two-kernel path (geoa3_kappa_loss_fwd, then geoa3_loss_bwd with the real upstream gradient): same losses, same
    gradient up to the rounding of (g*w)*x vs g*(w*x), for non-uniform upstream gradients, and bitwise reproducible."""
    from geoa3_b200 import loss_utils as L

    adv, ori, nrm = make(b, n, 3, std)
    Oc, Nr = cu(ori), cu(nrm)
    ko = L._get_kappa_ori(Oc, Nr, k)
    up = torch.linspace(0.3, 1.7, b, device="cuda")
    res = {}
    for fused in (True, False, True):
        L.FUSE_FWD_BWD = fused
        try:
            L.clear_cache()
            a = cu(adv).requires_grad_(True)
            tot, cd, hd, cv = L.geo_loss(a, Oc, Nr, ko, k, 1.0, 0.1, w_cu, single_side=single, hints=L.HintBuffers())
            (tot * up).sum().backward()
            cur = (tot.detach().clone(), cd.clone(), hd.clone(), cv.clone(), a.grad.clone())
        finally:
            L.FUSE_FWD_BWD = True
        if fused and True in res:
            assert all(torch.equal(x, y) for x, y in zip(cur, res[True])), "fused path is not reproducible"
        res[fused] = cur
    for x, y in zip(res[True][:4], res[False][:4]):
        assert float((x - y).abs().max()) <= 2e-6 * float(y.abs().max()) + 1e-12
    gf, gu = res[True][4], res[False][4]
    assert float((gf - gu).abs().max()) <= 2e-6 * float(gu.abs().max())
    # forward only (no gradient requested): the two-kernel forward runs, same values
    L.clear_cache()
    with torch.no_grad():
        t2 = L.geo_loss(cu(adv), Oc, Nr, ko, k, 1.0, 0.1, w_cu, single_side=single, hints=L.HintBuffers())[0]
    assert float((t2 - res[False][0]).abs().max()) <= 2e-6 * float(t2.abs().max())


def test_reference_api_cells_equal_scan():
    """The reference's own call pattern (chamfer_loss, hausdorff_loss, _get_kappa_adv, curvature_loss called one after
    the other on the same clouds, several steps in a row so that the second step is seeded by the first) with the
    cell-grid searches and with the scanning kernels: same values and gradients up to summation-order rounding."""
    from geoa3_b200 import loss_utils as L

    b, n, k = 3, 1024, 16
    adv0, ori, nrm = make(b, n, 5, 1e-2)
    Oc, Nr = cu(ori), cu(nrm)
    res = {}
    for cells in (True, False):
        L.USE_CELLS = cells
        try:
            L.clear_cache()
            ko = L._get_kappa_ori(Oc, Nr, k)
            cur, outs = adv0, []
            for step in range(3):
                cur = (cur + synth.make_offsets(b, n, seed=90 + step, std=4e-3)).astype(np.float32)
                a = cu(cur).requires_grad_(True)
                cd, hd = L.chamfer_loss(a, Oc), L.hausdorff_loss(a, Oc)
                kap, _ = L._get_kappa_adv(a, Oc, Nr, k)
                cv = L.curvature_loss(a, Oc, kap, ko)
                (cd + 0.1 * hd + cv).sum().backward()
                outs.append((cd.detach(), hd.detach(), cv.detach(), kap.detach(), a.grad.clone()))
            res[cells] = outs
        finally:
            L.USE_CELLS = True
    for s_c, s_s in zip(res[True], res[False]):
        assert torch.equal(s_c[0], s_s[0]) and torch.equal(s_c[1], s_s[1])          # CD / HD: identical 1-NN results
        for x, y in zip(s_c[2:], s_s[2:]):
            assert float((x - y).abs().max()) <= 2e-6 * float(y.abs().max()) + 1e-12


def test_full_batch_cells_parity_B250():
    """BASELINE config[1] size, EVERY cloud against the C oracle: cell-grid 1-NN (both directions, distances and
    argmins) and cell-grid kNN member sets (K = 17, hinted by the unperturbed cloud's lists), plus the size-independent
    properties: results do not depend on the hint or on the grid, and a second call on its own output is a fixed point."""
    from geoa3_b200 import ops

    b, n, k = 250, 1024, 16
    pc, _, _ = synth.make_batch(50, n, 0)
    ori = np.tile(pc, (5, 1, 1))
    adv = (ori + synth.make_offsets(b, n, seed=3, std=1e-2)).astype(np.float32)
    A, Oc = cu(adv), cu(ori)
    ca, co = ops.cell_sort(A, kref=k + 1), ops.cell_sort(Oc, kref=4)
    d1, j1, d2, i2 = ops.nn_pair_cells(ca, co)
    od1, oj1 = O.nn1(adv, ori)
    od2, oi2 = O.nn1(ori, adv)
    assert np.array_equal(j1.cpu().numpy(), oj1) and np.array_equal(i2.cpu().numpy(), oi2)
    assert np.array_equal(d1.cpu().numpy(), od1) and np.array_equal(d2.cpu().numpy(), od2)
    oi = O.knn(adv, adv, k + 1)[0][:, :, 1:]
    hint = cu(O.knn(ori, ori, k + 1)[0][:, :, 1:])   # "previous step" = the unperturbed cloud
    mem = ops.knn_cells(ca, k + 1, drop=1, hint=hint)[0]
    assert np.array_equal(np.sort(mem.cpu().numpy(), -1), np.sort(oi, -1))
    assert torch.equal(ops.knn_cells(ca, k + 1, drop=1, hint=mem)[0], mem)            # fixed point, same order
    assert torch.equal(ops.knn_cells(ca, k + 1, drop=1)[0], mem)                      # no hint: same members, same order
    other = ops.knn_cells(ops.cell_sort(A, grid=7), k + 1, drop=1, hint=hint)[0]      # another grid: same members
    assert torch.equal(other.sort(-1)[0], mem.sort(-1)[0])
    d1b, j1b, d2b, i2b = ops.nn_pair_cells(ops.cell_sort(A, grid=(5, 9, 3)), ops.cell_sort(Oc, grid=11), hint_a2o=i2[:, :n], hint_o2a=j1)
    assert torch.equal(j1b, j1) and torch.equal(i2b, i2) and torch.equal(d1b, d1) and torch.equal(d2b, d2)
