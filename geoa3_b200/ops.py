"""Thin tensor-level wrappers over the C ABI (one Python function per entry point of
include/geoa3_b200.h).  They allocate outputs with torch, launch on torch's current stream and raise
RuntimeError on any non-zero return code.  No autograd here — see loss_utils.py / pointnet2_ops."""
import math

import torch

from . import _lib
from ._lib import check, ptr, require_cuda_f32, require_cuda_i32, stream


LAUNCHES = 0  # own kernels launched through this module (bench.py reports it as gpu_launches)


def _count(n=1):
    global LAUNCHES
    LAUNCHES += n


def _guard(t):
    return torch.cuda.device(t.device)


def morton_order(pc, bits=10):
    """pc [b,3,n] -> (perm [b,n] int32, iperm [b,n] int32): points sorted by the Morton code of their coordinates
    quantised to `bits` bits per axis over the cloud's bounding box.  Plain torch ops — computed ONCE per attack
    on the original cloud, not on the per-step path.  Any permutation is valid for the kernels; this one makes
    groups of 32 consecutive points spatially compact so that bounding-box pruning bites."""
    with torch.no_grad():
        lo = pc.amin(2, keepdim=True)
        ext = (pc.amax(2, keepdim=True) - lo).amax(1, keepdim=True).clamp_min(1e-20)
        q = ((pc - lo) / ext * (2 ** bits - 1)).round().clamp_(0, 2 ** bits - 1).to(torch.int64)

        def spread(v):  # insert two zero bits between the bits of a 10-bit value
            v = (v | (v << 16)) & 0x030000FF
            v = (v | (v << 8)) & 0x0300F00F
            v = (v | (v << 4)) & 0x030C30C3
            v = (v | (v << 2)) & 0x09249249
            return v

        code = spread(q[:, 0]) | (spread(q[:, 1]) << 1) | (spread(q[:, 2]) << 2)
        perm = torch.sort(code, dim=1, stable=True)[1]
        iperm = torch.empty_like(perm)
        iperm.scatter_(1, perm, torch.arange(pc.shape[2], device=pc.device).expand_as(perm))
        return perm.to(torch.int32).contiguous(), iperm.to(torch.int32).contiguous()


def slab_order(pc):
    """pc [b,3,n] -> (perm, iperm) int32: points sorted along each cloud's WIDEST axis.  Groups of 32 consecutive
    points are then thin slabs, and a query's search ball of radius r only meets the slabs within +-r of it.
    Better than Morton boxes while r is a sizeable fraction of the cloud (k = 16 of n = 1024: r ~ 0.26 on a unit
    sphere, slab pruning keeps ~30 % of the candidates, Morton boxes ~65 %); Morton wins for n >= 4096."""
    with torch.no_grad():
        ext = pc.amax(2) - pc.amin(2)                                     # [b,3]
        axis = ext.argmax(1)                                              # [b]
        key = torch.gather(pc, 1, axis[:, None, None].expand(-1, 1, pc.shape[2]))[:, 0]
        perm = torch.sort(key, dim=1, stable=True)[1]
        iperm = torch.empty_like(perm)
        iperm.scatter_(1, perm, torch.arange(pc.shape[2], device=pc.device).expand_as(perm))
        return perm.to(torch.int32).contiguous(), iperm.to(torch.int32).contiguous()


def visit_order(pc):
    """Visiting order for the pruned searches: slabs along the widest axis for n < 2048, Morton above (measured on
    B200, B=250: nn_pair at n=1024 56 us with slabs vs 65 with Morton boxes; the Morton order is what makes the
    kNN scan scale at n >= 4096).  Any order is exact — this only decides how much the boxes prune."""
    return slab_order(pc) if pc.shape[2] < 2048 else morton_order(pc)


def arrange(pc, perm, with_bbox=False):
    """pc [b,3,n], perm [b,n] int32 -> pc with position t holding original point perm[t]; with_bbox also returns
    the group boxes of the arranged cloud (same launch; see geoa3_arrange in include/geoa3_b200.h)."""
    require_cuda_f32(pc, "pc"); require_cuda_i32(perm, "perm")
    b, c, n = pc.shape
    if c != 3 or tuple(perm.shape) != (b, n):
        raise RuntimeError("expected a [b,3,n] cloud and a [b,n] permutation")
    out = torch.empty_like(pc)
    bb = None
    if with_bbox:
        bb = torch.empty(b, _lib.load().geoa3_group_bbox_floats(n) // 8, 8, device=pc.device, dtype=torch.float32)
    with _guard(pc):
        _count(1)
        check(_lib.load().geoa3_arrange(ptr(pc), ptr(perm), b, n, ptr(out), ptr(bb), stream(pc)))
    return (out, bb) if with_bbox else out


def nn_pair(adv, ori, both=True, hint_a2o=None, hint_o2a=None, out=None, perm_a=None, perm_o=None, iperm_a=None,
            iperm_o=None, ori_arranged=None, adv_arranged=None):
    """adv [b,3,n], ori [b,3,m] -> d_a2o [b,n], jstar [b,n] i32, d_o2a [b,m] | None, istar [b,m] | None.
    hint_* (int32, optional) seed the search (exact for any seed); `out` = (d1, j1, d2, i2) preallocated
    buffers (j1/i2 may be the hint tensors themselves: in-place refresh of persistent hints).
    perm_* / iperm_*: visiting order for the pruned search; *_arranged: the clouds already in that order."""
    require_cuda_f32(adv, "adv_pc"); require_cuda_f32(ori, "ori_pc")
    b, c, n = adv.shape
    m = ori.shape[2]
    if c != 3 or ori.shape[0] != b or ori.shape[1] != 3:
        raise RuntimeError("expected [b,3,n] and [b,3,m] clouds")
    if out is not None:
        d1, j1, d2, i2 = out
    else:
        d1 = torch.empty(b, n, device=adv.device, dtype=torch.float32)
        j1 = torch.empty(b, n, device=adv.device, dtype=torch.int32)
        d2 = torch.empty(b, m, device=adv.device, dtype=torch.float32) if both else None
        i2 = torch.empty(b, m, device=adv.device, dtype=torch.int32) if both else None
    for h in (hint_a2o, hint_o2a, perm_a, perm_o, iperm_a, iperm_o):
        if h is not None:
            require_cuda_i32(h, "hint / perm")
    if (perm_a is None) != (perm_o is None):
        raise RuntimeError("perm_a and perm_o come together")
    if perm_a is not None:  # the kernel wants both clouds arranged in visiting order (coalesced staging)
        adv = adv_arranged if adv_arranged is not None else arrange(adv, perm_a)
        ori = ori_arranged if ori_arranged is not None else arrange(ori, perm_o)
    with _guard(adv):
        _count(1)
        check(_lib.load().geoa3_nn_pair(ptr(adv), ptr(ori), b, n, m, ptr(perm_a), ptr(perm_o), ptr(iperm_a),
                                        ptr(iperm_o), ptr(hint_a2o), ptr(hint_o2a) if both else None, ptr(d1),
                                        ptr(j1), ptr(d2), ptr(i2), stream(adv)))
    return d1, j1, d2, i2


def group_bbox(pc_arranged):
    """pc [b,3,n] already in visiting order -> boxes [b, G0+G1, 8] (see include/geoa3_b200.h)."""
    require_cuda_f32(pc_arranged, "pc_arranged")
    b, _, n = pc_arranged.shape
    nf = _lib.load().geoa3_group_bbox_floats(n)
    bb = torch.empty(b, nf // 8, 8, device=pc_arranged.device, dtype=torch.float32)
    with _guard(pc_arranged):
        _count(1)
        check(_lib.load().geoa3_group_bbox(ptr(pc_arranged), b, n, ptr(bb), stream(pc_arranged)))
    return bb


def knn(query, ref, K, drop=0, return_dist=False, hint=None, out=None, perm_q=None, perm_c=None, iperm_c=None,
        arranged=None, members_only=False):
    """query [b,3,n], ref [b,3,m] -> idx [b,n,K-drop] i32 (ascending (dist,idx)), dist | None.
    hint [b,n,hk] int32 (optional) only tightens the start threshold (exact for any hint); `out` may be the
    hint tensor itself (in-place refresh).  perm_* / iperm_c: visiting order for the pruned search;
    arranged = (cloud in that order, its boxes) from `arrange(..., with_bbox=True)` for a self-query whose
    arrangement already exists this step.
    members_only=True: the same K-drop members in ascending visiting order instead of by distance (geoa3_knn_set:
    for consumers that only sum over the neighbourhood; ~2x cheaper)."""
    require_cuda_f32(query, "query"); require_cuda_f32(ref, "ref")
    b, _, n = query.shape
    m = ref.shape[2]
    idx = out if out is not None else torch.empty(b, n, K - drop, device=query.device, dtype=torch.int32)
    bb = None
    if perm_c is not None:  # arrange the clouds in visiting order (coalesced staging) and box the ref groups
        same = query is ref and perm_q is perm_c
        if arranged is not None and same:
            ref, bb = arranged
        else:
            ref, bb = arrange(ref, perm_c, with_bbox=True)
        query = ref if same else (arrange(query, perm_q) if perm_q is not None else query)
    elif perm_q is not None:
        query = arrange(query, perm_q)
    hk = 0
    if hint is not None:
        require_cuda_i32(hint, "hint")
        hk = hint.shape[2]
    dist = torch.empty(b, n, K - drop, device=query.device, dtype=torch.float32) if return_dist else None
    with _guard(query):
        _count(1)
        fn = _lib.load().geoa3_knn_set if members_only else _lib.load().geoa3_knn
        check(fn(ptr(query), ptr(ref), b, n, m, K, drop, ptr(perm_q), ptr(perm_c), ptr(iperm_c), ptr(bb), ptr(hint), hk,
                 ptr(idx), ptr(dist), stream(query)))
    return idx, dist


class Cells(object):
    """Cell-grid blobs of a batch of clouds (geoa3_cell_sort): blobs uint8 [b, blob_bytes], n points, table capacity."""
    __slots__ = ("blobs", "n", "ncap")

    def __init__(self, blobs, n, ncap):
        self.blobs, self.n, self.ncap = blobs, n, ncap


def cell_table_capacity(n):
    """Default capacity of the cell table: room for ~2 cells per point (cells hold ~5 points on a surface; thin or
    clustered clouds need the slack), within what the sorting pass can keep in shared memory."""
    return int(max(64, min(2 * n, 8192, _lib.load().geoa3_cell_grid_max(n))))


def cell_sort(pc, kref=17.0, grid=None, ncap=None, out=None):
    """pc [b,3,n] -> Cells for `knn_cells` / `nn_pair_cells`.  grid=None: per-cloud cubic cells sized for balls of `kref`
    points (the search it will serve: K for kNN lists, a few points for 1-NN); grid=g or (gx,gy,gz): that grid.
    `out`: a Cells object of the same geometry to rebuild in place (CUDA-graph friendly)."""
    require_cuda_f32(pc, "pc")
    b, _, n = pc.shape
    gx, gy, gz = (0, 0, 0) if grid is None else ((int(grid),) * 3 if isinstance(grid, int) else tuple(int(g) for g in grid))
    if out is not None:
        ncap = out.ncap
    elif ncap is None:
        ncap = max(cell_table_capacity(n), gx * gy * gz)
    nb = _lib.load().geoa3_cell_blob_bytes(n, ncap)
    blobs = out.blobs if out is not None else torch.empty(b, nb, device=pc.device, dtype=torch.uint8)
    with _guard(pc):
        _count(1)
        check(_lib.load().geoa3_cell_sort(ptr(pc), b, n, ncap, float(kref), gx, gy, gz, ptr(blobs), stream(pc)))
    return out if out is not None else Cells(blobs, n, ncap)


def knn_cells(cells, K, drop=0, return_dist=False, hint=None, out=None):
    """Self-kNN member sets from the Cells of `cell_sort`: idx [b,n,K-drop] i32 in ORIGINAL numbering (rows by original
    query index, members in ascending cell-arrangement position), dist | None.  Exact for any hint / grid."""
    blobs, n = cells.blobs, cells.n
    b = blobs.shape[0]
    idx = out if out is not None else torch.empty(b, n, K - drop, device=blobs.device, dtype=torch.int32)
    hk = 0
    if hint is not None:
        require_cuda_i32(hint, "hint")
        hk = hint.shape[2]
    dist = torch.empty(b, n, K - drop, device=blobs.device, dtype=torch.float32) if return_dist else None
    with _guard(blobs):
        _count(1)
        check(_lib.load().geoa3_knn_cells(ptr(blobs), b, n, cells.ncap, K, drop, ptr(hint), hk, ptr(idx), ptr(dist),
                                          stream(blobs)))
    return idx, dist


def nn_pair_cells(cells_adv, cells_ori, both=True, hint_a2o=None, hint_o2a=None, out=None):
    """nn_pair over Cells (geoa3_nn_pair_cells): -> d_a2o [b,n], jstar [b,n], d_o2a [b,m] | None, istar | None."""
    n, m = cells_adv.n, cells_ori.n
    b, dev = cells_adv.blobs.shape[0], cells_adv.blobs.device
    if out is not None:
        d1, j1, d2, i2 = out
    else:
        d1 = torch.empty(b, n, device=dev, dtype=torch.float32)
        j1 = torch.empty(b, n, device=dev, dtype=torch.int32)
        d2 = torch.empty(b, m, device=dev, dtype=torch.float32) if both else None
        i2 = torch.empty(b, m, device=dev, dtype=torch.int32) if both else None
    for h, nm in ((hint_a2o, "hint_a2o"), (hint_o2a, "hint_o2a")):
        if h is not None:
            require_cuda_i32(h, nm)
    with _guard(cells_adv.blobs):
        _count(1)
        check(_lib.load().geoa3_nn_pair_cells(ptr(cells_adv.blobs), ptr(cells_ori.blobs), b, n, m, cells_adv.ncap,
                                              cells_ori.ncap, ptr(hint_a2o), ptr(hint_o2a), ptr(d1), ptr(j1), ptr(d2),
                                              ptr(i2), stream(cells_adv.blobs)))
    return d1, j1, d2, i2


def kappa_loss_fwd(pc, normal=None, jstar=None, nbr=None, d_a2o=None, d_o2a=None, kappa_ori=None, m=None,
                   want_kappa=True, want_nrm=False, want_cd=False, want_hd=False, want_curv=False):
    """One launch: kappa / borrowed normals / per-cloud CD, HD(+argmax), curvature loss. Returns a dict."""
    require_cuda_f32(pc, "pc")
    b, _, n = pc.shape
    if m is None:
        m = normal.shape[2] if normal is not None else (d_o2a.shape[1] if d_o2a is not None else n)
    k = nbr.shape[2] if nbr is not None else 0
    dev = pc.device
    out = {}
    kap = torch.empty(b, n, device=dev, dtype=torch.float32) if (want_kappa and k > 0) else None
    nrm = torch.empty(b, 3, n, device=dev, dtype=torch.float32) if (want_nrm and k > 0) else None
    cd = torch.empty(b, device=dev, dtype=torch.float32) if want_cd else None
    hd = torch.empty(b, device=dev, dtype=torch.float32) if want_hd else None
    ha = torch.empty(b, device=dev, dtype=torch.int32) if want_hd else None
    cu = torch.empty(b, device=dev, dtype=torch.float32) if (want_curv and k > 0) else None
    with _guard(pc):
        _count(1)
        check(_lib.load().geoa3_kappa_loss_fwd(ptr(pc), ptr(normal), ptr(jstar), ptr(nbr), k, ptr(d_a2o), ptr(d_o2a),
                                               ptr(kappa_ori), b, n, m, ptr(kap), ptr(nrm), ptr(cd), ptr(hd), ptr(ha),
                                               ptr(cu), stream(pc)))
    out.update(kappa=kap, nrm=nrm, cd=cd, hd=hd, hd_arg=ha, curv=cu)
    return out


def loss_bwd(adv, ori=None, nrm_adv=None, kappa_adv=None, kappa_ori=None, jstar=None, istar=None, nbr=None,
             hd_arg=None, g_cd=None, g_hd=None, g_cu=None, g_kappa=None, m=None):
    """Fused deterministic backward -> grad_adv [b,3,n]."""
    require_cuda_f32(adv, "adv_pc")
    b, _, n = adv.shape
    if m is None:
        m = ori.shape[2] if ori is not None else n
    k = nbr.shape[2] if nbr is not None else 0
    grad = torch.empty_like(adv)
    for g in (g_cd, g_hd, g_cu, g_kappa):
        if g is not None:
            require_cuda_f32(g, "upstream gradient")
    lib = _lib.load()
    wsb = int(lib.geoa3_loss_bwd_workspace_bytes(b, n, m, k))  # 0 unless the lists outgrow shared memory
    ws = torch.empty(wsb, device=adv.device, dtype=torch.uint8) if wsb else None
    with _guard(adv):
        _count(1 if not wsb else 4)
        check(lib.geoa3_loss_bwd(ptr(adv), ptr(ori), ptr(nrm_adv), ptr(kappa_adv), ptr(kappa_ori), ptr(jstar),
                                 ptr(istar), ptr(nbr), ptr(hd_arg), ptr(g_cd), ptr(g_hd), ptr(g_cu),
                                 ptr(g_kappa), b, n, m, k, ptr(grad), ptr(ws), wsb, stream(adv)))
    return grad


def geo_fwd_bwd_supported(n, m, k):
    return bool(_lib.load().geoa3_geo_fwd_bwd_supported(n, m, k))


def geo_fwd_bwd(adv, ori, normal, kappa_ori, jstar, istar, nbr, d_a2o, d_o2a, w_cd, w_hd, w_cu):
    """Fused kappa / CD / HD / curvature forward + unit-upstream backward (geoa3_geo_fwd_bwd) ->
    dict(cd, hd, curv [b], grad [b,3,n])."""
    require_cuda_f32(adv, "adv_pc")
    b, _, n = adv.shape
    m = ori.shape[2]
    k = nbr.shape[2] if nbr is not None else 0
    dev = adv.device
    out = {"cd": torch.empty(b, device=dev), "hd": torch.empty(b, device=dev), "curv": torch.empty(b, device=dev),
           "grad": torch.empty_like(adv)}
    with _guard(adv):
        _count(1)
        check(_lib.load().geoa3_geo_fwd_bwd(ptr(adv), ptr(ori), ptr(normal), ptr(kappa_ori), ptr(jstar), ptr(istar),
                                            ptr(nbr), k, ptr(d_a2o), ptr(d_o2a), float(w_cd), float(w_hd), float(w_cu),
                                            b, n, m, ptr(out["cd"]), ptr(out["hd"]), ptr(out["curv"]), None, None, None,
                                            ptr(out["grad"]), stream(adv)))
    return out


def farthest_points_sample_idx(points, nsamples, start):
    """points [b,n,3] f32, start [b] int32 -> idx [b,nsamples] int32: plain FPS from the given first picks
    (geoa3_farthest_points_sample; Lib/utility.py:175-187 semantics)."""
    require_cuda_f32(points, "points"); require_cuda_i32(start, "start")
    b, n, _ = points.shape
    out = torch.empty(b, nsamples, device=points.device, dtype=torch.int32)
    with _guard(points):
        _count(1)
        check(_lib.load().geoa3_farthest_points_sample(ptr(points), b, n, nsamples, ptr(start), ptr(out), stream(points)))
    return out


# ------------------------------------------------------------------ pointnet2_ops (bindings.cpp:6-19 names)
def furthest_point_sampling(points, nsamples):
    require_cuda_f32(points, "points")
    b, n, _ = points.shape
    out = torch.empty(b, nsamples, device=points.device, dtype=torch.int32)
    with _guard(points):
        _count(1)
        check(_lib.load().geoa3_furthest_point_sampling(ptr(points), b, n, nsamples, ptr(out), stream(points)))
    return out


def gather_points(points, idx):
    require_cuda_f32(points, "points"); require_cuda_i32(idx, "idx")
    b, c, n = points.shape
    m = idx.shape[1]
    out = torch.empty(b, c, m, device=points.device, dtype=torch.float32)
    with _guard(points):
        _count(1)
        check(_lib.load().geoa3_gather_points(ptr(points), ptr(idx), b, c, n, m, ptr(out), stream(points)))
    return out


def _workspace(dev, b, n, npoints, nsample):
    nbytes = _lib.load().geoa3_group_points_grad_workspace_bytes(b, n, npoints, nsample)
    return torch.empty(nbytes, device=dev, dtype=torch.uint8), nbytes


def gather_points_grad(grad_out, idx, n):
    require_cuda_f32(grad_out, "grad_out"); require_cuda_i32(idx, "idx")
    b, c, m = grad_out.shape
    out = torch.empty(b, c, n, device=grad_out.device, dtype=torch.float32)
    ws, nbytes = _workspace(grad_out.device, b, n, m, 1)
    with _guard(grad_out):
        _count(2)
        check(_lib.load().geoa3_gather_points_grad(ptr(grad_out), ptr(idx), b, c, n, m, ptr(out), ptr(ws), nbytes,
                                                   stream(grad_out)))
    return out


def ball_query(new_xyz, xyz, radius, nsample):
    require_cuda_f32(new_xyz, "new_xyz"); require_cuda_f32(xyz, "xyz")
    b, n, _ = xyz.shape
    m = new_xyz.shape[1]
    idx = torch.empty(b, m, nsample, device=xyz.device, dtype=torch.int32)
    with _guard(xyz):
        _count(1)
        check(_lib.load().geoa3_ball_query(ptr(new_xyz), ptr(xyz), b, n, m, float(radius), int(nsample), ptr(idx),
                                           stream(xyz)))
    return idx


def group_points(points, idx):
    require_cuda_f32(points, "points"); require_cuda_i32(idx, "idx")
    b, c, n = points.shape
    _, npoints, nsample = idx.shape
    out = torch.empty(b, c, npoints, nsample, device=points.device, dtype=torch.float32)
    with _guard(points):
        _count(1)
        check(_lib.load().geoa3_group_points(ptr(points), ptr(idx), b, c, n, npoints, nsample, ptr(out), stream(points)))
    return out


def group_points_grad(grad_out, idx, n):
    require_cuda_f32(grad_out, "grad_out"); require_cuda_i32(idx, "idx")
    b, c, npoints, nsample = grad_out.shape
    out = torch.empty(b, c, n, device=grad_out.device, dtype=torch.float32)
    ws, nbytes = _workspace(grad_out.device, b, n, npoints, nsample)
    with _guard(grad_out):
        _count(2)
        check(_lib.load().geoa3_group_points_grad(ptr(grad_out), ptr(idx), b, c, n, npoints, nsample, ptr(out), ptr(ws),
                                                  nbytes, stream(grad_out)))
    return out


def three_nn(unknowns, knows):
    require_cuda_f32(unknowns, "unknowns"); require_cuda_f32(knows, "knows")
    b, n, _ = unknowns.shape
    m = knows.shape[1]
    dist2 = torch.empty(b, n, 3, device=unknowns.device, dtype=torch.float32)
    idx = torch.empty(b, n, 3, device=unknowns.device, dtype=torch.int32)
    with _guard(unknowns):
        _count(1)
        check(_lib.load().geoa3_three_nn(ptr(unknowns), ptr(knows), b, n, m, ptr(dist2), ptr(idx), stream(unknowns)))
    return dist2, idx


def three_interpolate(points, idx, weight):
    require_cuda_f32(points, "points"); require_cuda_i32(idx, "idx"); require_cuda_f32(weight, "weight")
    b, c, m = points.shape
    n = idx.shape[1]
    out = torch.empty(b, c, n, device=points.device, dtype=torch.float32)
    with _guard(points):
        _count(1)
        check(_lib.load().geoa3_three_interpolate(ptr(points), ptr(idx), ptr(weight), b, c, m, n, ptr(out),
                                                  stream(points)))
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    require_cuda_f32(grad_out, "grad_out"); require_cuda_i32(idx, "idx"); require_cuda_f32(weight, "weight")
    b, c, n = grad_out.shape
    out = torch.empty(b, c, m, device=grad_out.device, dtype=torch.float32)
    ws, nbytes = _workspace(grad_out.device, b, m, n, 3)
    with _guard(grad_out):
        _count(2)
        check(_lib.load().geoa3_three_interpolate_grad(ptr(grad_out), ptr(idx), ptr(weight), b, c, n, m, ptr(out),
                                                       ptr(ws), nbytes, stream(grad_out)))
    return out
