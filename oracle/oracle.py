"""ctypes/numpy front-end of oracle/geoa3_oracle.c — TEST INFRASTRUCTURE ONLY.

Every function takes/returns numpy arrays; see geoa3_oracle.c for the reference file:line each
routine restates.  Index routines are bit-exact fp32; value routines return float64.
"""
import ctypes as C
import os
import os.path as osp
import subprocess

import numpy as np

_HERE = osp.dirname(osp.abspath(__file__))
_SO = osp.join(_HERE, "libgeoa3_oracle.so")
_lib = None

f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")


def build(force=False):
    src = osp.join(_HERE, "geoa3_oracle.c")
    if force or not osp.exists(_SO) or osp.getmtime(_SO) < osp.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libgeoa3_oracle.so"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _vp(a):
    return a.ctypes.data_as(C.c_void_p)


def pairdist(Q, R):
    Q, R = _f32(Q), _f32(R)
    b, _, n = Q.shape
    m = R.shape[2]
    out = np.empty((b, n, m), np.float32)
    lib().orc_pairdist(_vp(Q), _vp(R), b, n, m, _vp(out))
    return out


def nn1(Q, R):
    """Q [b,3,n], R [b,3,m] -> (dmin [b,n] f32, arg [b,n] i32), ties -> lowest index."""
    Q, R = _f32(Q), _f32(R)
    b, _, n = Q.shape
    m = R.shape[2]
    d = np.empty((b, n), np.float32)
    a = np.empty((b, n), np.int32)
    lib().orc_nn1(_vp(Q), _vp(R), b, n, m, _vp(d), _vp(a))
    return d, a


def knn(Q, R, K):
    """K smallest (dist, idx) lexicographic, ascending -> (idx [b,n,K] i32, dist [b,n,K] f32)."""
    Q, R = _f32(Q), _f32(R)
    b, _, n = Q.shape
    m = R.shape[2]
    idx = np.empty((b, n, K), np.int32)
    dist = np.empty((b, n, K), np.float32)
    lib().orc_knn(_vp(Q), _vp(R), b, n, m, K, _vp(idx), _vp(dist))
    return idx, dist


def kappa(P, nrm, nbr):
    P, nrm, nbr = _f32(P), _f32(nrm), _i32(nbr)
    b, _, n = P.shape
    k = nbr.shape[2]
    out = np.empty((b, n), np.float64)
    lib().orc_kappa(_vp(P), _vp(nrm), _vp(nbr), b, n, k, _vp(out))
    return out


def gather3(src, idx):
    src, idx = _f32(src), _i32(idx)
    b, _, ns = src.shape
    no = idx.shape[1]
    out = np.empty((b, 3, no), np.float32)
    lib().orc_gather3(_vp(src), _vp(idx), b, ns, no, _vp(out))
    return out


def kappa_ori(pc, normal, k):
    """loss_utils.py:52-62"""
    idx, _ = knn(pc, pc, k + 1)
    nbr = np.ascontiguousarray(idx[:, :, 1:])
    return kappa(pc, normal, nbr), nbr


def geo_forward(adv, ori, normal, kap_ori, k):
    """Everything the fused loss group produces for one step. Returns a dict."""
    adv, ori, normal = _f32(adv), _f32(ori), _f32(normal)
    b, _, n = adv.shape
    d1, jstar = nn1(adv, ori)
    d2, istar = nn1(ori, adv)
    out = dict(d_a2o=d1, jstar=jstar, d_o2a=d2, istar=istar)
    if k > 0:
        idx, _ = knn(adv, adv, k + 1)
        nbr = np.ascontiguousarray(idx[:, :, 1:])
        nrm_adv = gather3(normal, jstar)
        kap = kappa(adv, nrm_adv, nbr)
        out.update(nbr=nbr, nrm_adv=nrm_adv, kappa_adv=kap)
    else:
        kap = None
    cd = np.empty(b, np.float64)
    hd = np.empty(b, np.float64)
    cu = np.empty(b, np.float64)
    ha = np.empty(b, np.int32)
    ko = _f32(kap_ori) if kap_ori is not None else np.zeros((b, n), np.float32)
    lib().orc_loss_fwd(_vp(d1), _vp(d2), _vp(jstar), _vp(kap) if kap is not None else None, _vp(ko), b, n,
                       _vp(cd), _vp(hd), _vp(ha), _vp(cu))
    out.update(cd=cd, hd=hd, hd_arg=ha, curv=cu)
    return out


def geo_backward(adv, ori, fwd, kap_ori, g_cd, g_hd, g_cu):
    """Closed-form d(g_cd*CD + g_hd*HD + g_cu*CUR)/d adv, float64 [b,3,n]."""
    adv, ori = _f32(adv), _f32(ori)
    b, _, n = adv.shape
    has_k = "nbr" in fwd
    k = fwd["nbr"].shape[2] if has_k else 0
    nbr = _i32(fwd["nbr"]) if has_k else np.zeros((1,), np.int32)
    nrm_adv = _f32(fwd["nrm_adv"]) if has_k else np.zeros((b, 3, n), np.float32)
    ko = _f32(kap_ori) if kap_ori is not None else np.zeros((b, n), np.float32)
    G = np.empty((b, 3, n), np.float64)
    g_cd, g_hd, g_cu = (_f64(np.broadcast_to(np.asarray(x, np.float64), (b,))) for x in (g_cd, g_hd, g_cu))
    lib().orc_loss_bwd(_vp(adv), _vp(ori), _vp(nrm_adv), _vp(ko), _vp(_i32(fwd["jstar"])), _vp(_i32(fwd["istar"])),
                       _vp(nbr), _vp(_f32(fwd["d_a2o"])), _vp(g_cd), _vp(g_hd), _vp(g_cu), b, n, k, _vp(G))
    return G


# ------------------------------------------------------------------ neighbourhood regularisers
def _nbr_vec(P, nbr):
    """P [b,3,n] f32, nbr [b,n,k] -> P[nbr] - P as float64 [b,3,n,k]."""
    P = np.asarray(P, np.float64)
    b, _, n = P.shape
    out = np.empty((b, 3, n, nbr.shape[2]), np.float64)
    for i in range(b):
        out[i] = P[i][:, nbr[i]] - P[i][:, :, None]
    return out


def aux_losses(adv, ori, normal, k, h=0.03, threshold_coef=1.05):
    """Values (float64 tails on fp32-exact neighbour lists) of the reference's regularisers,
    Lib/loss_utils.py:99-149: displacement, corresponding-normal, repulsion, distance-kmean, kNN smoothing."""
    adv32, ori32 = _f32(adv), _f32(ori)
    nbr_o = knn(ori32, ori32, k + 1)[0][:, :, 1:]
    nbr_a = knn(adv32, adv32, k + 1)[0][:, :, 1:]
    a, o, nr = (np.asarray(x, np.float64) for x in (adv32, ori32, normal))
    b, _, n = a.shape
    theta = ((a - o) ** 2).sum(1)                                                        # :105
    nn_theta = np.stack([theta[i][nbr_o[i]] for i in range(b)])
    out = dict(displacement=((nn_theta - theta[:, :, None]) ** 2).mean(2))               # :107
    v = _nbr_vec(adv32, nbr_a)
    L = np.maximum(np.sqrt((v ** 2).sum(1, keepdims=True)), 1e-12)
    out["corr_normal"] = np.abs(((v / L) * nr[:, :, :, None]).sum(1)).mean(2)            # :115-117
    dis = (v ** 2).sum(1)
    out["repulsion"] = -(dis * np.exp(-(dis ** 2) / (h ** 2))).mean(2)                   # :123
    dk = np.sqrt(((1e-12 - v) ** 2).sum(1))                                              # :127 (p_i - p_j + 1e-12)
    dm = dk.mean(-1)
    dmk = np.stack([dm[i][nbr_a[i]] for i in range(b)])
    out["kmean"] = np.abs(dm[:, :, None] - dmk).mean(-1)                                 # :133
    kd = dis.mean(-1)
    thr = kd.mean(-1) + threshold_coef * kd.std(-1, ddof=1)                              # torch.std is unbiased
    out["smoothing"] = (kd * (kd > thr[:, None])).mean(1)                                # :146-149
    return out


def uniform_loss(adv, percentages=(0.004, 0.006, 0.008, 0.010, 0.012), radius=1.0, k=2):
    """Lib/loss_utils.py:151-190 with the pointnet2 calls it intends (fps / ball_query / group restated above)."""
    import math

    adv32 = _f32(adv)
    b, _, n = adv32.shape
    xyz = np.ascontiguousarray(adv32.transpose(0, 2, 1))
    npoint = int(n * 0.05)
    seeds = fps(xyz, npoint)
    new_xyz = np.stack([xyz[i][seeds[i]] for i in range(b)])
    total = 0.0
    for p in percentages:
        p = p * 4
        nsample = int(n * p)
        r = math.sqrt(p * radius)
        expect_len = float(np.sqrt(np.float32(math.pi * (radius ** 2) * p / nsample)))   # fp32 tensor at :160-161
        idx = ball_query(new_xyz, xyz, r, nsample)
        patches = np.stack([xyz[i][idx[i]] for i in range(b)]).reshape(b * npoint, nsample, 3)
        pt = np.ascontiguousarray(patches.transpose(0, 2, 1))
        nb = knn(pt, pt, k + 1)[0][:, :, 1:]
        d = (_nbr_vec(pt, nb) ** 2).sum(1)
        u = np.sqrt(np.abs(d) + 1e-12).mean(-1)
        u = (u - expect_len) ** 2 / (expect_len + 1e-12)
        total += u.mean() * math.pow(p * 100, 2)
    return total / len(percentages)


# ------------------------------------------------------------------ pointnet2_ops
def opt_n_threads(w):
    return lib().orc_opt_n_threads(int(w))


def fps(xyz, m):
    """xyz [b,n,3] -> idx [b,m] i32 (sampling_gpu.cu:69-173)."""
    xyz = _f32(xyz)
    b, n, _ = xyz.shape
    out = np.zeros((b, m), np.int32)
    lib().orc_fps(_vp(xyz), b, n, m, _vp(out))
    return out


def fps_from(xyz, m, start):
    """xyz [b,n,3], start [b] -> idx [b,m] i32: plain FPS from given start indices (Lib/utility.py:175-187)."""
    xyz, start = _f32(xyz), _i32(start)
    b, n, _ = xyz.shape
    out = np.zeros((b, m), np.int32)
    lib().orc_fps_from(_vp(xyz), b, n, m, _vp(start), _vp(out))
    return out


def ball_query(new_xyz, xyz, radius, nsample):
    new_xyz, xyz = _f32(new_xyz), _f32(xyz)
    b, n, _ = xyz.shape
    m = new_xyz.shape[1]
    out = np.zeros((b, m, nsample), np.int32)
    lib().orc_ball_query(_vp(new_xyz), _vp(xyz), b, n, m, C.c_float(radius), nsample, _vp(out))
    return out


def group_points(points, idx):
    points, idx = _f32(points), _i32(idx)
    b, c, n = points.shape
    _, np_, ns = idx.shape
    out = np.empty((b, c, np_, ns), np.float32)
    lib().orc_group_points(_vp(points), _vp(idx), b, c, n, np_, ns, _vp(out))
    return out


def group_points_grad(grad_out, idx, n):
    grad_out, idx = _f32(grad_out), _i32(idx)
    b, c, np_, ns = grad_out.shape
    out = np.empty((b, c, n), np.float64)
    lib().orc_group_points_grad(_vp(grad_out), _vp(idx), b, c, n, np_, ns, _vp(out))
    return out


def gather_points(points, idx):
    points, idx = _f32(points), _i32(idx)
    b, c, n = points.shape
    m = idx.shape[1]
    out = np.empty((b, c, m), np.float32)
    lib().orc_gather_points(_vp(points), _vp(idx), b, c, n, m, _vp(out))
    return out


def gather_points_grad(grad_out, idx, n):
    grad_out, idx = _f32(grad_out), _i32(idx)
    b, c, m = grad_out.shape
    out = np.empty((b, c, n), np.float64)
    lib().orc_gather_points_grad(_vp(grad_out), _vp(idx), b, c, n, m, _vp(out))
    return out


def three_nn(unknown, known):
    unknown, known = _f32(unknown), _f32(known)
    b, n, _ = unknown.shape
    m = known.shape[1]
    d = np.empty((b, n, 3), np.float32)
    i = np.empty((b, n, 3), np.int32)
    lib().orc_three_nn(_vp(unknown), _vp(known), b, n, m, _vp(d), _vp(i))
    return d, i


def three_interpolate(points, idx, weight):
    points, idx, weight = _f32(points), _i32(idx), _f32(weight)
    b, c, m = points.shape
    n = idx.shape[1]
    out = np.empty((b, c, n), np.float64)
    lib().orc_three_interpolate(_vp(points), _vp(idx), _vp(weight), b, c, m, n, _vp(out))
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    grad_out, idx, weight = _f32(grad_out), _i32(idx), _f32(weight)
    b, c, n = grad_out.shape
    out = np.empty((b, c, m), np.float64)
    lib().orc_three_interpolate_grad(_vp(grad_out), _vp(idx), _vp(weight), b, c, n, m, _vp(out))
    return out
