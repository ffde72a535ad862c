"""Generates tests/golden/loss_*.npz by running the REFERENCE's own Lib/loss_utils.py (verbatim, from
/root/reference, with its un-vendored pytorch3d dependency stubbed as documented in
oracle/ref_loader.py).  Run in the build container only:  python tests/golden/make_golden.py

Each fixture holds fp32 inputs (adv, ori, normal), the reference's fp32 outputs and autograd gradient of
sum_b(CD + 0.1*HD + CUR) (weights of main_attack.py:342-349), plus an fp64 run of the same reference code
for tight value checks.  Index tensors are not stored: the reference API does not return them; index parity
is pinned by oracle/geoa3_oracle.c (see its header)."""
import os.path as osp
import sys

import numpy as np
import torch

ROOT = osp.dirname(osp.dirname(osp.dirname(osp.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader, synth  # noqa: E402

HERE = osp.dirname(osp.abspath(__file__))
CASES = [
    # name, b, n, k, offset std, first instance
    ("loss_b3_n256_k16", 3, 256, 16, 2e-2, 0),
    ("loss_b2_n500_k8", 2, 500, 8, 1e-3, 3),     # ragged n, reference-sized initial perturbation
    ("loss_b4_n128_k16_big", 4, 128, 16, 1e-1, 6),  # large perturbation: many-to-one argmins, empty columns
]
NORMAL_CASES = [(3, 512, 16), (2, 300, 8)]   # b, n, k
FPS_CASES = [(3, 2048, 256), (2, 1000, 100), (1, 5000, 64), (2, 300, 300)]   # b, n, picks
DEFENSE_CASES = [  # n, outlier_knn, alpha, drop_num, offset std
    (1024, 2, 1.1, 50, 1e-2),
    (700, 8, 0.5, 123, 3e-2),
    (333, 16, 2.0, 1, 1e-3),
]
AUX_CASES = [
    ("aux_b2_n512_k4", 2, 512, 4, 1e-2, 1),
    ("aux_b3_n300_k8", 3, 300, 8, 3e-2, 5),
]


def main(which=("loss", "l2", "aux", "defense", "fps", "normal")):
    torch.manual_seed(0)
    torch.set_num_threads(4)
    for name, b, n, k, std, start in (CASES if "loss" in which else []):
        pc, nr, lab = synth.make_batch(b, n, start)
        adv = pc + synth.make_offsets(b, n, seed=7 + start, std=std)
        r32 = ref_loader.geo_loss_and_grad(adv, pc, nr, k, dtype=torch.float32)
        r64 = ref_loader.geo_loss_and_grad(adv, pc, nr, k, dtype=torch.float64)
        out = dict(adv=adv, ori=pc, normal=nr, k=np.int32(k))
        for key, v in r32.items():
            out["f32_" + key] = v
        for key, v in r64.items():
            out["f64_" + key] = v
        np.savez_compressed(osp.join(HERE, name + ".npz"), **out)
        print(name, "cd", r32["cd"], "hd", r32["hd"], "curv", r32["curv"])

    # norm_l2_loss (Lib/loss_utils.py:25-26): value and autograd gradient of the reference function, fp32 + fp64
    if "l2" in which:
        pc, _, _ = synth.make_batch(3, 200, 2)
        adv = pc + synth.make_offsets(3, 200, seed=5, std=3e-2)
        out = dict(adv=adv, ori=pc)
        for tag, dt in (("f32_", torch.float32), ("f64_", torch.float64)):
            a = torch.from_numpy(adv).to(dt).requires_grad_(True)
            v = ref_loader.load().norm_l2_loss(a, torch.from_numpy(pc).to(dt))
            v.sum().backward()
            out[tag + "l2"], out[tag + "grad"] = v.detach().numpy(), a.grad.numpy()
        np.savez_compressed(osp.join(HERE, "l2_case.npz"), **out)
        print("l2_case", out["f32_l2"])

    # neighbourhood regularisers (Lib/loss_utils.py:99-190), same reference module, k neighbours
    for name, b, n, k, std, start in (AUX_CASES if "aux" in which else []):
        pc, nr, lab = synth.make_batch(b, n, start)
        adv = pc + synth.make_offsets(b, n, seed=11 + start, std=std)
        out = dict(adv=adv, ori=pc, normal=nr, k=np.int32(k))
        for tag, dt in (("f32_", torch.float32), ("f64_", torch.float64)):
            for key, v in ref_loader.aux_losses_and_grads(adv, pc, nr, k, dtype=dt).items():
                out[tag + key] = v
        np.savez_compressed(osp.join(HERE, name + ".npz"), **out)
        print(name, {key: float(np.asarray(out["f32_" + key]).mean()) for key in
                     ("displacement", "corr_normal", "repulsion", "kmean", "smoothing", "uniform")})

    # defense.py statistical outlier filters (:25-40) on adversarially perturbed single clouds with planted outliers
    if "defense" in which:
        out = {}
        for i, (n, knn, alpha, drop, std) in enumerate(DEFENSE_CASES):
            pc, _, _ = synth.make_batch(1, n, i)
            adv = pc + synth.make_offsets(1, n, seed=21 + i, std=std)
            adv[0, :, ::37] += synth.make_offsets(1, n, seed=31 + i, std=0.15)[0, :, ::37]   # a few far outliers
            out["c%d_pc" % i] = adv.astype(np.float32)
            out["c%d_args" % i] = np.array([drop, alpha, knn], np.float64)
            for key, v in ref_loader.defense_outputs(adv, drop, alpha, knn).items():
                out["c%d_%s" % (i, key)] = v
            print("defense case", i, "removed", out["c%d_var_num" % i], out["c%d_fix_num" % i])
        np.savez_compressed(osp.join(HERE, "defense_cases.npz"), **out)

    # farthest_points_sample (Lib/utility.py:175-187) with fixed first picks
    if "fps" in which:
        out = {}
        for i, (b, n, m) in enumerate(FPS_CASES):
            pc, _, _ = synth.make_batch(b, n, 3 * i)
            start = np.random.default_rng(i).integers(0, n, b).astype(np.int32)
            out["c%d_pc" % i], out["c%d_start" % i] = pc, start
            out["c%d_sel" % i] = ref_loader.ref_farthest_points_sample(pc, m, start)
            print("fps case", i, out["c%d_sel" % i].shape)
        np.savez_compressed(osp.join(HERE, "fps_plain_cases.npz"), **out)

    # estimate_normal (Lib/utility.py:40-90) — neighbourhood PCA normals
    if "normal" in which:
        out = {}
        for i, (b, n, k) in enumerate(NORMAL_CASES):
            pc, _, _ = synth.make_batch(b, n, 2 * i)
            out["c%d_pc" % i], out["c%d_k" % i] = pc, np.int32(k)
            out["c%d_normal" % i] = ref_loader.ref_estimate_normal(pc, k)
            print("normal case", i, out["c%d_normal" % i].shape)
        np.savez_compressed(osp.join(HERE, "estimate_normal_cases.npz"), **out)


if __name__ == "__main__":
    main(tuple(sys.argv[1:]) or ("loss", "l2", "aux", "defense", "fps", "normal"))   # e.g. `make_golden.py aux` regenerates only aux_*
