"""GPU tests of the attack driver (the caller of the hot path): graph replay == eager, sharded == unsharded,
reference return convention."""
import numpy as np
import pytest
import torch

from oracle import synth

pytestmark = pytest.mark.gpu


def _data(b, n, start=0):
    pc, nr, lab = synth.make_batch(b, n, start)
    return [torch.from_numpy(pc).permute(0, 2, 1)[:, None].contiguous(), torch.from_numpy(nr).permute(0, 2, 1)[:, None].contiguous(),
            torch.from_numpy(lab)[:, None]]


def _net():
    from geoa3_b200.victims import build_victim

    torch.manual_seed(0)
    return build_victim("PointNet").cuda().eval()


def test_attack_return_convention_and_graph_equals_eager():
    from geoa3_b200 import attack as atk

    net = _net()
    cfg = atk.make_cfg(binary_max_steps=2, iter_max_steps=6, curv_loss_knn=8)
    data = _data(3, 256)
    out_g = atk.attack(net, data, cfg, use_cuda_graph=True)
    out_e = atk.attack(net, data, cfg, use_cuda_graph=False)
    best, target, success, steps, losses = out_g
    assert best.shape == (3, 3, 256) and target.shape == (3,) and success.dtype == np.bool_ and len(steps) == 3
    assert len(losses) == 6 and len(losses[0]) == 3
    assert np.allclose(np.asarray(losses), np.asarray(out_e[4]), rtol=1e-4, atol=1e-5)
    assert list(success) == list(out_e[2])


def test_sharded_equals_unsharded():
    from geoa3_b200 import attack as atk
    from geoa3_b200 import dist as gd

    net = _net()
    cfg = atk.make_cfg(binary_max_steps=1, iter_max_steps=5, curv_loss_knn=8)
    data = _data(4, 256)
    full = atk.attack(net, data, cfg, use_cuda_graph=False)
    parts = []
    for r in range(2):
        rows = gd.shard_rows(4, 2, r)
        sl = [d[rows.start:rows.stop] for d in data]
        parts.append(atk.attack(net, sl, cfg, use_cuda_graph=False, global_batch=4, rows=list(rows)))
    cat = np.concatenate([np.asarray(p[4]) for p in parts], 1)
    assert np.allclose(cat, np.asarray(full[4]), rtol=1e-4, atol=1e-5)


def test_projection_helpers():
    from geoa3_b200 import attack as atk
    from oracle import oracle as O

    pc, nr, _ = synth.make_batch(2, 200, 2)
    off = synth.make_offsets(2, 200, std=5e-2)
    P, N_, F_ = torch.from_numpy(pc).cuda(), torch.from_numpy(nr).cuda(), torch.from_numpy(off).cuda()
    real = atk.find_offset(P, P + F_)
    _, j = O.nn1(pc + off, pc)
    exp = (pc + off) - np.take_along_axis(pc, j[:, None, :].repeat(3, 1), 2)
    assert np.allclose(real.cpu().numpy(), exp, atol=1e-7)
    proj = atk.offset_proj(F_, P, N_)
    _, j2 = O.nn1(off, pc)  # the reference queries with the raw offset (geoA3_attack.py:65)
    nn_ = np.take_along_axis(nr, j2[:, None, :].repeat(3, 1), 2)
    u = nn_ / (np.sqrt((nn_ ** 2).sum(1, keepdims=True)) + 1e-6)
    assert np.allclose(proj.cpu().numpy(), (off * u).sum(1, keepdims=True) * u, atol=1e-6)
    clip = atk.lp_clip(F_, 0.01)
    assert float((clip ** 2).sum(1).sqrt().max()) <= 0.01 * (1 + 1e-5)


def test_pointnetpp_attack_step_runs():
    from geoa3_b200 import attack as atk
    from geoa3_b200.victims import build_victim

    torch.manual_seed(0)
    net = build_victim("PointNetPP_ssg").cuda().eval()
    cfg = atk.make_cfg(binary_max_steps=1, iter_max_steps=3)
    out = atk.attack(net, _data(2, 1024), cfg, use_cuda_graph=True)
    assert np.isfinite(np.asarray(out[4])).all()


def test_attack_step_with_uniform_loss_captures():
    """--uniform_loss_weight != 0 (geoA3_attack.py:169-171): the step still captures into a CUDA graph and the
    replay reproduces the eager run."""
    from geoa3_b200 import attack as atk

    net = _net()
    cfg = atk.make_cfg(binary_max_steps=1, iter_max_steps=4, curv_loss_knn=8, uniform_loss_weight=0.5)
    data = _data(3, 512)
    out_g = atk.attack(net, data, cfg, use_cuda_graph=True)
    out_e = atk.attack(net, data, cfg, use_cuda_graph=False)
    assert np.allclose(np.asarray(out_g[4]), np.asarray(out_e[4]), rtol=1e-4, atol=1e-5)
    cfg0 = atk.make_cfg(binary_max_steps=1, iter_max_steps=4, curv_loss_knn=8)
    assert not np.allclose(np.asarray(out_e[4]), np.asarray(atk.attack(net, data, cfg0, use_cuda_graph=False)[4]))


def test_attack_trajectory_matches_cpu_port():
    """Several full attack iterations (victim forward, CE + CD + 0.1 HD + curvature, backward, Adam) on the GPU
    driver vs the dense CPU port of the reference step (oracle/torch_port.py, pinned against the reference's loss
    functions by tests/test_port_cpu.py): same weights, same initial offsets => the per-instance loss trajectory
    agrees.  TF32 is switched off for the victim so that the comparison is fp32 against fp32."""
    from geoa3_b200 import attack as atk
    from oracle import torch_port as P

    tf32 = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        net = _net()
        b, n, k, steps = 3, 256, 8, 6
        pc, nr, lab = synth.make_batch(b, n, 0)
        cfg = atk.make_cfg(binary_max_steps=1, iter_max_steps=steps, curv_loss_knn=k)
        dev = torch.device("cuda")
        pc_t, nr_t, lab_t = torch.from_numpy(pc).to(dev), torch.from_numpy(nr).to(dev), torch.from_numpy(lab).to(dev)
        st = atk.AttackState(net, pc_t, nr_t, lab_t, lab_t, cfg, targeted=False, global_batch=b)
        st.begin_search_step(0, atk.default_offsets(b, n, 0, 0).to(dev))
        for _ in range(steps):
            st.run_step()
        got = st.loss_log[:steps].cpu().numpy()

        import copy
        cpu_net = copy.deepcopy(net).cpu().eval()
        ref = P.CpuAttackStep(cpu_net, torch.from_numpy(pc), torch.from_numpy(nr), torch.from_numpy(lab), k=k, seed=0)
        want = np.stack([ref.step(True).numpy() for _ in range(steps)])
        assert np.allclose(got[0], want[0], rtol=2e-5, atol=1e-6)          # first step: pure forward parity
        assert np.allclose(got, want, rtol=2e-3, atol=1e-4), np.abs(got - want).max()   # Adam amplifies rounding
        assert np.abs(got[-1] - got[0]).max() > 1e-3                         # the optimisation actually moved
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32


@pytest.mark.parametrize("n,k", [(4096, 16), (10000, 32)])
def test_attack_runs_at_sweep_sizes(n, k):
    """BASELINE config[3] cloud sizes through the whole attack step (pruned searches, large-cloud backward, CUDA
    graph): graph replay == eager, finite losses."""
    from geoa3_b200 import attack as atk

    net = _net()
    cfg = atk.make_cfg(binary_max_steps=1, iter_max_steps=3, curv_loss_knn=k)
    data = _data(2, n)
    out_g = atk.attack(net, data, cfg, use_cuda_graph=True)
    out_e = atk.attack(net, data, cfg, use_cuda_graph=False)
    lg, le = np.asarray(out_g[4]), np.asarray(out_e[4])
    assert np.isfinite(lg).all() and np.allclose(lg, le, rtol=1e-4, atol=1e-5)


def test_attack_subsample_opt():
    """--is_subsample_opt (geoA3_attack.py:283-296): a 2048-point cloud attacked through a 512-point victim input;
    every step draws fresh farthest-point subsamples; the full cloud is what gets optimised and returned."""
    from geoa3_b200 import attack as atk

    net = _net()
    cfg = atk.make_cfg(binary_max_steps=1, iter_max_steps=8, curv_loss_knn=8, npoint=512, is_subsample_opt=True, eval_num=3)
    data = _data(2, 2048)
    torch.manual_seed(1)
    best, target, success, steps, losses = atk.attack(net, data, cfg)
    L = np.asarray(losses)
    assert best.shape == (2, 3, 2048) and L.shape == (8, 2) and np.isfinite(L).all()
    assert not np.allclose(L[0], L[-1])


def test_attack_partial_variable():
    """--is_partial_var (geoA3_attack.py:239-262,279-280): only the knn_range neighbours of a random seed point move
    in each 50-step period; earlier periods stay frozen in the accumulated cloud."""
    from geoa3_b200 import attack as atk

    net = _net()
    b, n, kr = 3, 256, 3
    cfg = atk.make_cfg(binary_max_steps=1, iter_max_steps=60, curv_loss_knn=8, is_partial_var=True, knn_range=kr)
    pc, nr, lab = synth.make_batch(b, n, 0)
    dev = torch.device("cuda")
    pc_t, nr_t, lab_t = torch.from_numpy(pc).to(dev), torch.from_numpy(nr).to(dev), torch.from_numpy(lab).to(dev)
    st = atk.AttackState(net, pc_t, nr_t, lab_t, lab_t, cfg, targeted=False, global_batch=b)
    np.random.seed(3)
    st.begin_search_step(0, torch.zeros(b, 3, n, device=dev))
    moved_after = {}
    for i in range(60):
        st.run_step()
        if i in (49, 59):
            moved_after[i] = ((st.base + st.offset - pc_t).abs().sum(1) > 0).sum(1).cpu().numpy()
    assert (moved_after[49] == kr).all()                                  # first period: exactly the region
    assert ((moved_after[59] >= kr) & (moved_after[59] <= 2 * kr)).all()  # second period adds a second region
    assert np.isfinite(st.loss_log[:60].cpu().numpy()).all()
    # the public entry point accepts the flag
    out = atk.attack(net, _data(2, 256), atk.make_cfg(binary_max_steps=1, iter_max_steps=5, curv_loss_knn=8,
                                                      is_partial_var=True))
    assert out[0].shape == (2, 3, 256)


def test_attack_pre_jitter_input():
    """--is_pre_jitter_input: tangent-plane jitter re-drawn every few steps, success judged on the clean cloud."""
    from geoa3_b200 import attack as atk

    net = _net()
    cfg = atk.make_cfg(binary_max_steps=1, iter_max_steps=7, curv_loss_knn=8, is_pre_jitter_input=True,
                       calculate_project_jitter_noise_iter=3, jitter_k=8)
    torch.manual_seed(2)
    out = atk.attack(net, _data(3, 256), cfg)
    L = np.asarray(out[4])
    assert out[0].shape == (3, 3, 256) and L.shape == (7, 3) and np.isfinite(L).all()


@pytest.mark.parametrize("kw", [
    dict(cls_loss_type="Margin", attack_label="All", confidence=0.5),
    dict(dis_loss_type="L2", hd_loss_weight=0.0, curv_loss_weight=0.0),
    dict(is_pro_grad=True, is_real_offset=True, cc_linf=0.02),
    dict(optim="sgd", is_use_lr_scheduler=True),
    dict(is_cd_single_side=True, cls_loss_type="None"),
], ids=["margin_targeted", "l2_only", "projected_clipped", "sgd_scheduler", "single_side_no_cls"])
def test_attack_flag_matrix(kw):
    """Every remaining flag combination of the point-cloud attack loop runs to finite losses (and, where the step is
    graph-capturable, the replay equals the eager run)."""
    from geoa3_b200 import attack as atk

    net = _net()
    cfg = atk.make_cfg(binary_max_steps=2, iter_max_steps=4, curv_loss_knn=8, **kw)
    data = _data(3, 256)
    if kw.get("attack_label") == "All":
        data = data + [torch.tensor([[5], [7], [9]])]
    out = atk.attack(net, data, cfg, use_cuda_graph=True)
    L = np.asarray(out[4])
    assert L.shape == (4, 3) and np.isfinite(L).all()
    if not kw.get("is_use_lr_scheduler"):
        Le = np.asarray(atk.attack(net, data, cfg, use_cuda_graph=False)[4])
        assert np.allclose(L, Le, rtol=1e-4, atol=1e-5)
    if "cc_linf" in kw:
        assert float((out[0] - torch.from_numpy(synth.make_batch(3, 256)[0]).cuda()).abs().max()) <= 0.02 + 1e-6 \
            or not out[2].any()


def test_fold_batchnorm_and_frozen_parameters():
    """attack() runs on a BatchNorm-folded copy of the eval-mode victim with its parameters frozen (no weight-gradient
    kernels); the caller's module is left exactly as it was.  Folding is exact algebra: logits and the first-step
    losses agree with the unfolded victim to rounding."""
    from geoa3_b200 import attack as atk
    from geoa3_b200.victims import build_victim, fold_batchnorm

    torch.manual_seed(3)
    for arch, n in (("PointNet", 512), ("PointNetPP_ssg", 1024)):
        net = build_victim(arch).cuda().eval()
        for m in net.modules():  # non-trivial running statistics
            if isinstance(m, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d)):
                m.running_mean.normal_(0, 0.1); m.running_var.uniform_(0.5, 1.5)
                m.weight.data.uniform_(0.5, 1.5); m.bias.data.normal_(0, 0.1)
        x = torch.from_numpy(synth.make_batch(3, n, 1)[0]).cuda()
        folded = fold_batchnorm(net, x[:2])
        assert not any(isinstance(m, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d)) for m in folded.modules())
        tf32 = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
        torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
        try:
            with torch.no_grad():
                a, b_ = net(x), folded(x)
            assert float((a - b_).abs().max()) <= 2e-5 * max(1.0, float(a.abs().max()))
            before = {k: v.clone() for k, v in net.state_dict().items()}
            data = _data(3, n)
            cfg = atk.make_cfg(binary_max_steps=1, iter_max_steps=2, curv_loss_knn=8)
            l_fold = np.asarray(atk.attack(net, data, cfg, use_cuda_graph=False, fold_bn=True)[4])
            l_plain = np.asarray(atk.attack(net, data, cfg, use_cuda_graph=False, fold_bn=False)[4])
            assert np.allclose(l_fold[0], l_plain[0], rtol=2e-5, atol=1e-6)
        finally:
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
        assert all(p.requires_grad for p in net.parameters())           # restored
        assert all(p.grad is None for p in net.parameters())            # and never differentiated
        after = net.state_dict()
        assert list(after) == list(before) and all(torch.equal(after[k], before[k]) for k in before)
