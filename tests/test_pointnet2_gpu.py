"""GPU parity (-m gpu) of the pointnet2_ops drop-in against the CPU oracle and, when the prebuilt
oracle/_ref/pointnet2_ref_ext.so travelled to the box, against the reference's own CUDA kernels."""
import numpy as np
import pytest
import torch

from helpers import rel_err
from oracle import oracle as O
from oracle import synth

pytestmark = pytest.mark.gpu


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def clouds(b, n, start=0):
    pc, _, _ = synth.make_batch(b, n, start)
    return np.ascontiguousarray(pc.transpose(0, 2, 1))  # (b,n,3)


def ref_ext():
    from oracle import build_ref

    return build_ref.load_ref()


@pytest.mark.parametrize("b,n,m", [(10, 1024, 512), (10, 512, 128), (3, 1000, 100), (2, 37, 37), (2, 2048, 64),
                                   (1, 5000, 32)])
def test_fps_bitexact(b, n, m):
    from geoa3_b200.pointnet2_ops import pointnet2_utils as pu

    xyz = clouds(b, n)  # family 6 (flat plate) has points inside the |p|^2<=1e-3 skip zone
    got = pu.furthest_point_sample(cu(xyz), m)
    assert got.dtype == torch.int32 and not got.requires_grad
    assert np.array_equal(got.cpu().numpy(), O.fps(xyz, m))
    ext = ref_ext()
    if ext is not None:
        assert torch.equal(got, ext.furthest_point_sampling(cu(xyz), m))


def test_fps_ties_and_degenerate():
    from geoa3_b200.pointnet2_ops import pointnet2_utils as pu

    lat, _ = synth.lattice_cloud(216)
    xyz = np.ascontiguousarray(lat.T[None])
    zeros = np.full((1, 64, 3), 1e-3, np.float32)
    for x, m in ((xyz, 100), (zeros, 10), (np.tile(xyz, (1, 3, 1)), 50)):
        got = pu.furthest_point_sample(cu(x), m).cpu().numpy()
        assert np.array_equal(got, O.fps(x, m))
        ext = ref_ext()
        if ext is not None:
            assert np.array_equal(got, ext.furthest_point_sampling(cu(x), m).cpu().numpy())


@pytest.mark.parametrize("b,n,m,r,ns", [(10, 1024, 512, 0.2, 64), (10, 512, 128, 0.4, 64), (4, 1024, 512, 0.1, 16),
                                        (4, 1024, 512, 0.4, 128), (2, 999, 77, 0.3, 33), (2, 100, 10, 0.01, 8),
                                        (2, 10000, 500, 0.44, 480), (1, 6000, 64, 2.5, 300)])   # nsample > 256: uniform_loss on dense clouds
def test_ball_query_bitexact(b, n, m, r, ns):
    from geoa3_b200.pointnet2_ops import pointnet2_utils as pu

    xyz = clouds(b, n, 1)
    new = np.ascontiguousarray(xyz[:, O.fps(xyz, m)[0], :]) if b == 1 else np.stack(
        [xyz[i, O.fps(xyz[i:i + 1], m)[0]] for i in range(b)])
    new[:, -1] += 7.0  # a centroid with no neighbour at all -> all-zero row
    got = pu.ball_query(r, ns, cu(xyz), cu(new))
    assert got.dtype == torch.int32 and got.shape == (b, m, ns)
    assert np.array_equal(got.cpu().numpy(), O.ball_query(new, xyz, r, ns))
    ext = ref_ext()
    if ext is not None:
        assert torch.equal(got, ext.ball_query(cu(new), cu(xyz), r, ns))


def _shell(rng, n, centre, r):
    """n points at distance r (+- a few ulp) from `centre`: every ball / arg-max decision is ulp-critical"""
    u = rng.standard_normal((n, 3))
    u /= np.linalg.norm(u, axis=1, keepdims=True)
    return (np.asarray(centre, np.float64) + r * u).astype(np.float32)


def test_ball_query_ulp_critical_vs_reference_binary():
    """Points on the ball's surface: membership flips with the last ulp of d2, so this only passes when the
    distance arithmetic (contraction order y,x,z) matches the reference binary exactly."""
    from geoa3_b200.pointnet2_ops import pointnet2_utils as pu

    rng = np.random.default_rng(5)
    b, n, m, ns = 4, 2048, 16, 256
    r = np.float32(0.3)
    new = rng.uniform(-0.5, 0.5, (b, m, 3)).astype(np.float32)
    xyz = np.stack([np.concatenate([_shell(rng, n // m, new[i, j], float(r)) for j in range(m)]) for i in range(b)])
    got = pu.ball_query(float(r), ns, cu(xyz), cu(new))
    assert np.array_equal(got.cpu().numpy(), O.ball_query(new, xyz, float(r), ns))
    inside = (got.cpu().numpy() != got.cpu().numpy()[:, :, :1]).sum()
    assert inside > 1000  # the shell really straddles the boundary (neither all in nor all out)
    ext = ref_ext()
    if ext is not None:
        assert torch.equal(got, ext.ball_query(cu(new), cu(xyz), float(r), ns))


def test_fps_ulp_critical_vs_reference_binary():
    """All points (nearly) equidistant from the start point and from each other's images: arg-max decided by ulps."""
    from geoa3_b200.pointnet2_ops import pointnet2_utils as pu

    rng = np.random.default_rng(6)
    xs = []
    for i in range(6):
        c = rng.uniform(0.3, 0.6, 3)
        pts = _shell(rng, 1023, c, 0.25)
        xs.append(np.concatenate([c[None].astype(np.float32), pts]))
    xyz = np.stack(xs)
    got = pu.furthest_point_sample(cu(xyz), 256)
    assert np.array_equal(got.cpu().numpy(), O.fps(xyz, 256))
    ext = ref_ext()
    if ext is not None:
        assert torch.equal(got, ext.furthest_point_sampling(cu(xyz), 256))


@pytest.mark.parametrize("b,c,n,m,ns,r", [(4, 3, 1024, 512, 64, 0.3), (4, 128, 512, 128, 64, 0.3),
                                          (2, 67, 300, 50, 7, 0.3), (2, 320, 512, 128, 128, 0.3),
                                          # row-structured backward: every slices-per-row / channels-per-CTA
                                          # variant, odd row counts, channel tails, mostly-padding rows
                                          (2, 9, 512, 50, 32, 0.3), (2, 16, 1024, 37, 32, 0.12),
                                          (2, 10, 300, 33, 64, 0.08), (2, 6, 400, 21, 128, 0.5),
                                          (2, 13, 2048, 19, 64, 0.2),
                                          (2, 5, 301, 20, 8, 0.3),    # odd n: rows not 16-byte sized => plain staging
                                          (1, 24, 10000, 40, 32, 0.15)])  # long rows: fewer channels staged per CTA
def test_group_points_and_grad(b, c, n, m, ns, r):
    from geoa3_b200.pointnet2_ops import pointnet2_utils as pu

    rng = np.random.default_rng(0)
    xyz = clouds(b, n, 2)
    new = np.stack([xyz[i, O.fps(xyz[i:i + 1], m)[0]] for i in range(b)])
    idx = O.ball_query(new, xyz, r, ns)
    feats = rng.standard_normal((b, c, n)).astype(np.float32)
    f = cu(feats).requires_grad_(True)
    out = pu.grouping_operation(f, cu(idx))
    assert np.array_equal(out.detach().cpu().numpy(), O.group_points(feats, idx))
    go = rng.standard_normal(out.shape).astype(np.float32)
    out.backward(cu(go))
    ref = O.group_points_grad(go, idx, n)
    assert rel_err(f.grad.cpu().numpy(), ref) < 1e-5
    # deterministic: bitwise equal on a second run
    f2 = cu(feats).requires_grad_(True)
    pu.grouping_operation(f2, cu(idx)).backward(cu(go))
    assert torch.equal(f.grad, f2.grad)
    ext = ref_ext()
    if ext is not None:
        assert torch.equal(out.detach(), ext.group_points(cu(feats), cu(idx)))
        assert rel_err(f.grad.cpu().numpy(), ext.group_points_grad(cu(go), cu(idx), n).cpu().numpy()) < 1e-5


def test_group_grad_arbitrary_idx_takes_generic_path():
    """Index rows that do NOT have the ball-query shape (random, with repeats anywhere): the device-side shape
    check must route the call to the generic CSR path; results still match the oracle and stay deterministic."""
    from geoa3_b200.pointnet2_ops import pointnet2_utils as pu

    rng = np.random.default_rng(3)
    for (b, c, n, m, ns) in ((3, 16, 300, 40, 24), (2, 5, 128, 64, 32)):
        idx = rng.integers(0, n, (b, m, ns)).astype(np.int32)
        idx[:, ::3, 5:9] = idx[:, ::3, 1:2]  # repeats in the middle of a row
        feats = rng.standard_normal((b, c, n)).astype(np.float32)
        go = rng.standard_normal((b, c, m, ns)).astype(np.float32)
        grads = []
        for _ in range(2):
            f = cu(feats).requires_grad_(True)
            pu.grouping_operation(f, cu(idx)).backward(cu(go))
            grads.append(f.grad.clone())
        assert rel_err(grads[0].cpu().numpy(), O.group_points_grad(go, idx, n)) < 1e-5
        assert torch.equal(grads[0], grads[1])
        # one bad row among ball-query rows also forces the generic path
        xyz = clouds(b, n, 2)
        new = np.stack([xyz[i, O.fps(xyz[i:i + 1], m)[0]] for i in range(b)])
        idx2 = O.ball_query(new, xyz, 0.3, ns)
        idx2[0, 0, 1] = idx2[0, 0, 0]
        idx2[0, 0, 2] = (idx2[0, 0, 0] + 1) % n
        f = cu(feats).requires_grad_(True)
        pu.grouping_operation(f, cu(idx2)).backward(cu(go))
        assert rel_err(f.grad.cpu().numpy(), O.group_points_grad(go, idx2, n)) < 1e-5


def test_gather_and_grad():
    from geoa3_b200.pointnet2_ops import pointnet2_utils as pu

    rng = np.random.default_rng(1)
    b, c, n, m = 5, 3, 1024, 512
    feats = rng.standard_normal((b, c, n)).astype(np.float32)
    idx = rng.integers(0, n, (b, m)).astype(np.int32)
    idx[:, :5] = 7  # duplicates
    f = cu(feats).requires_grad_(True)
    out = pu.gather_operation(f, cu(idx))
    assert np.array_equal(out.detach().cpu().numpy(), O.gather_points(feats, idx))
    go = rng.standard_normal(out.shape).astype(np.float32)
    out.backward(cu(go))
    assert rel_err(f.grad.cpu().numpy(), O.gather_points_grad(go, idx, n)) < 1e-6


def test_three_nn_interpolate():
    from geoa3_b200.pointnet2_ops import pointnet2_utils as pu

    rng = np.random.default_rng(2)
    b, n, m, c = 3, 700, 190, 19
    unknown = clouds(b, n, 4)
    known = clouds(b, m, 5)
    dist, idx = pu.three_nn(cu(unknown), cu(known))
    od, oi = O.three_nn(unknown, known)
    assert np.array_equal(idx.cpu().numpy(), oi)
    assert np.array_equal(dist.cpu().numpy(), np.sqrt(od))
    w = rng.uniform(size=(b, n, 3)).astype(np.float32)
    pts = rng.standard_normal((b, c, m)).astype(np.float32)
    p = cu(pts).requires_grad_(True)
    out = pu.three_interpolate(p, idx, cu(w))
    assert rel_err(out.detach().cpu().numpy(), O.three_interpolate(pts, oi, w)) < 1e-6
    go = rng.standard_normal(out.shape).astype(np.float32)
    out.backward(cu(go))
    assert rel_err(p.grad.cpu().numpy(), O.three_interpolate_grad(go, oi, w, m)) < 1e-5
    ext = ref_ext()
    if ext is not None:
        rd, ri = ext.three_nn(cu(unknown), cu(known))
        assert torch.equal(ri, idx) and torch.equal(torch.sqrt(rd), dist)
        assert torch.equal(out.detach(), ext.three_interpolate(cu(pts), idx, cu(w)))  # bit-exact vs the binary
        assert rel_err(p.grad.cpu().numpy(), ext.three_interpolate_grad(cu(go), idx, cu(w), m).cpu().numpy()) < 1e-5


def test_query_and_group_module():
    from geoa3_b200.pointnet2_ops import pointnet2_utils as pu

    xyz = clouds(2, 256, 3)
    X = cu(xyz)
    fidx = pu.furthest_point_sample(X, 32)
    new_xyz = pu.gather_operation(X.transpose(1, 2).contiguous(), fidx).transpose(1, 2).contiguous()
    feats = torch.randn(2, 5, 256, device="cuda")
    out = pu.QueryAndGroup(0.3, 16, use_xyz=True)(X, new_xyz, feats)
    assert out.shape == (2, 8, 32, 16)
    idx = O.ball_query(new_xyz.cpu().numpy(), xyz, 0.3, 16)
    gx = O.group_points(np.ascontiguousarray(xyz.transpose(0, 2, 1)), idx) - new_xyz.cpu().numpy().transpose(0, 2, 1)[..., None]
    assert np.allclose(out[:, :3].cpu().numpy(), gx, atol=0, rtol=0)
    assert pu.GroupAll()(X, None, feats).shape == (2, 8, 1, 256)


def test_group_grad_row_shape_check_edges():
    """Rows at the edge of the ball-query shape: padding that starts exactly on a 32-entry boundary (valid), a stray
    index after the padding began, an equal neighbour pair, an index >= n in the ascending part (all invalid => the
    generic path must take over).  Whatever path runs, the gradient must equal the oracle's."""
    from geoa3_b200.pointnet2_ops import pointnet2_utils as pu

    rng = np.random.default_rng(5)
    b, c, n, m, ns = 2, 12, 400, 24, 64
    base = np.sort(rng.choice(n - 1, (b, m, ns), replace=True), axis=2).astype(np.int32)
    for i in range(b):
        for j in range(m):  # strictly ascending rows
            base[i, j] = np.sort(rng.choice(n - 1, ns, replace=False))
    feats = rng.standard_normal((b, c, n)).astype(np.float32)
    go = rng.standard_normal((b, c, m, ns)).astype(np.float32)

    def check(idx):
        f = cu(feats).requires_grad_(True)
        pu.grouping_operation(f, cu(idx)).backward(cu(go))
        assert rel_err(f.grad.cpu().numpy(), O.group_points_grad(go, idx, n)) < 1e-5

    v = base.copy(); v[:, :, 32:] = v[:, :, :1]; check(v)                      # valid: padding from position 32
    v = base.copy(); v[:, ::2, 17:] = v[:, ::2, :1]; check(v)                  # valid: padding from position 17
    w = v.copy(); w[0, 0, 40] = w[0, 0, 5]; check(w)                           # stray index inside the padding
    w = base.copy(); w[1, 3, 50] = w[1, 3, 49]; check(w)                       # equal neighbours, not the first index
    w = base.copy(); w[1, 3, 33] = w[1, 3, 31]; check(w)                       # descending across the chunk boundary
    w = v.copy(); w[0, 2, 63] = (w[0, 2, 0] + 1) % n; check(w)                 # last entry leaves the padding


def test_plain_fps_vs_oracle_and_reference_golden():
    """geoa3_farthest_points_sample (Lib/utility.py:175-187 semantics): indices bit-exact vs the C oracle for every
    kernel variant (n = 37 ... 5000), the reference function's own picks (golden), and gradient through the gather."""
    import os.path as osp

    from geoa3_b200 import ops, utility
    from helpers import GOLDEN_DIR

    rng = np.random.default_rng(9)
    for (b, n, m) in ((3, 37, 20), (4, 256, 256), (3, 500, 77), (5, 1024, 512), (2, 2048, 100), (2, 4096, 300),
                      (1, 10000, 128)):
        xyz = clouds(b, n, 1)
        start = rng.integers(0, n, b).astype(np.int32)
        got = ops.farthest_points_sample_idx(cu(xyz), m, cu(start)).cpu().numpy()
        assert np.array_equal(got, O.fps_from(xyz, m, start)), (b, n, m)
    g = np.load(osp.join(GOLDEN_DIR, "fps_plain_cases.npz"))
    i = 0
    while "c%d_pc" % i in g:
        pc, start, sel = g["c%d_pc" % i], g["c%d_start" % i], g["c%d_sel" % i]
        p = cu(pc).requires_grad_(True)
        out = utility.farthest_points_sample(p, sel.shape[2], start=cu(start))
        assert np.array_equal(out.detach().cpu().numpy(), sel)
        out.sum().backward()
        assert float(p.grad.sum()) == pc.shape[0] * 3 * sel.shape[2]  # one unit per selected coordinate
        i += 1
    # random first picks: valid, distinct indices
    r = utility.farthest_points_sample(cu(pc), 10)
    assert r.shape == (pc.shape[0], 3, 10)
