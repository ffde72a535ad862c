"""Data formats either side of the path (CPU): the .mat dataset layout and result files, checked against the
reference's own Provider class when /root/reference is present (same items, same tensors)."""
import importlib.util
import os.path as osp

import numpy as np
import pytest
import torch
from scipy.io import loadmat

from geoa3_b200 import provider, synth

REF = "/root/reference/Provider/modelnet10_instance250.py"


@pytest.fixture(scope="module")
def mat(tmp_path_factory):
    return provider.write_synthetic_mat(str(tmp_path_factory.mktemp("d") / "modelnet10_250instances_256.mat"), 250, 256)


def test_mat_layout_and_items(mat):
    raw = loadmat(mat)
    assert raw["data"].shape == (250, 3, 256) and raw["normal"].shape == (250, 3, 256) and raw["label"].shape == (250, 1)
    assert list(raw["label"][::25, 0]) == synth.CLASS_IDS           # class-major, 25 per class
    ds = provider.ModelNet40(mat, "All")
    pcs, nrs, gts, tgt = ds[30]
    assert len(ds) == 250 and pcs.shape == (9, 256, 3) and nrs.shape == (9, 256, 3) and gts.shape == (9,)
    assert int(gts[0]) == synth.CLASS_IDS[1] and int(gts[0]) not in tgt.tolist() and len(set(tgt.tolist())) == 9
    un = provider.ModelNet40(mat, "Untarget")[7]
    assert len(un) == 3 and un[0].shape == (1, 256, 3) and un[2].shape == (1,)
    chair = provider.ModelNet40(mat, "chair")
    assert len(chair) == 25 and chair.start_index == 100 and int(chair[0][2][0]) == 3
    half = provider.ModelNet40(mat, "All", is_half_forward=True)[0]
    assert half[0][0].shape[0] == 4 and half[1][0].shape[0] == 5
    rs = provider.ModelNet40(mat, "Untarget", resample_num=64)
    p = rs[3][0][0].numpy()
    assert p.shape == (64, 3) and abs(np.linalg.norm(p, axis=1).max() - 1) < 1e-6 and np.abs(p.mean(0)).max() < 1e-6


@pytest.mark.skipif(not osp.isfile(REF), reason="reference not present")
@pytest.mark.parametrize("mode", ["All", "Untarget", "sofa"])
def test_items_equal_reference_provider(mat, mode):
    spec = importlib.util.spec_from_file_location("ref_provider", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    a, b = provider.ModelNet40(mat, mode), ref.ModelNet40(mat, mode)
    assert len(a) == len(b) and a.start_index == b.start_index
    for i in (0, len(a) // 2, len(a) - 1):
        for x, y in zip(a[i], b[i]):
            assert torch.equal(x, y)


def test_result_files_roundtrip(tmp_path):
    pc, nr, lab = synth.make_instance(4, 128)
    name = provider.result_name(12, 3, 17, 17)
    provider.save_adversarial(str(tmp_path), name, pc, 3, 17, est_normal=nr)
    rec = loadmat(str(tmp_path / "Mat" / (name + ".mat")))
    assert np.array_equal(rec["adversary_point_clouds"], pc) and int(rec["gt_label"].item()) == 3 and int(rec["attack_label"].item()) == 17
    assert np.array_equal(rec["est_normal"], nr)
    rows = open(str(tmp_path / "PC" / (name + ".obj"))).read().splitlines()
    assert len(rows) == 128 and rows[0].startswith("v ") and rows[0].endswith(" 0 0 0")
    assert np.allclose([float(t) for t in rows[5].split()[1:4]], pc[:, 5], atol=1e-6)
