"""Pins the dense CPU port (oracle/torch_port.py, the cpu_baseline arm) against the reference golden vectors
and, when /root/reference is present, against the reference code executed verbatim."""
import numpy as np
import pytest
import torch

from helpers import golden_files, rel_err
from oracle import ref_loader, torch_port as P


@pytest.mark.parametrize("path", golden_files()[:2])
def test_port_matches_golden(path):
    g = np.load(path)
    k = int(g["k"])
    adv = torch.from_numpy(g["adv"]).double().requires_grad_(True)
    ori, nrm = torch.from_numpy(g["ori"]).double(), torch.from_numpy(g["normal"]).double()
    ko = P.get_kappa_ori(ori, nrm, k)
    con, cd, hd, cu = P.constrain_loss(adv, ori, nrm, ko, k, 1.0, 0.1, 1.0)
    con.sum().backward()
    assert rel_err(cd.detach().numpy(), g["f64_cd"]) < 1e-6
    assert rel_err(hd.detach().numpy(), g["f64_hd"]) < 1e-6
    assert rel_err(cu.detach().numpy(), g["f64_curv"]) < 1e-6
    assert rel_err(adv.grad.numpy(), g["f64_grad"]) < 1e-6


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present")
def test_port_matches_live_reference_fp32():
    from oracle import synth

    pc, nr, _ = synth.make_batch(2, 128, 4)
    adv = pc + synth.make_offsets(2, 128, seed=2, std=3e-2)
    ref = ref_loader.geo_loss_and_grad(adv, pc, nr, 8, dtype=torch.float32)
    a = torch.from_numpy(adv).requires_grad_(True)
    o, n_ = torch.from_numpy(pc), torch.from_numpy(nr)
    con, cd, hd, cu = P.constrain_loss(a, o, n_, P.get_kappa_ori(o, n_, 8), 8, 1.0, 0.1, 1.0)
    con.sum().backward()
    assert rel_err(a.grad.numpy(), ref["grad"]) < 1e-4
    assert rel_err(cd.detach().numpy(), ref["cd"]) < 1e-5


def test_cpu_attack_step_runs():
    from geoa3_b200.victims import PointNet
    from oracle import synth

    torch.manual_seed(0)
    net = PointNet(40).eval()
    pc, nr, lab = synth.make_batch(1, 128)
    st = P.CpuAttackStep(net, torch.from_numpy(pc), torch.from_numpy(nr), torch.from_numpy(lab), k=8)
    l0 = st.step()
    l1 = st.step()
    assert l0.shape == (1,) and torch.isfinite(l1).all()
