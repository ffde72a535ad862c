"""Drop-in for the reference's Lib/loss_utils.py geometry-aware losses, backed by libgeoa3_b200.so.

Same names, positional signatures, shapes and return conventions as the reference functions
(file:line under /root/reference):

    norm_l2_loss(adv_pc, ori_pc)                               Lib/loss_utils.py:25-26
    chamfer_loss(adv_pc, ori_pc)              -> [b]           Lib/loss_utils.py:28-35
    pseudo_chamfer_loss(adv_pc, ori_pc)       -> [b]           Lib/loss_utils.py:37-43
    hausdorff_loss(adv_pc, ori_pc)            -> [b]           Lib/loss_utils.py:45-50
    _get_kappa_ori(pc, normal, k=2)           -> [b,n]         Lib/loss_utils.py:52-62
    _get_kappa_adv(adv_pc, ori_pc, ori_normal, k=2) -> ([b,n], [b,3,n])   Lib/loss_utils.py:64-82
    curvature_loss(adv_pc, ori_pc, adv_kappa, ori_kappa, k=2) -> [b]      Lib/loss_utils.py:84-97

    displacement_loss(adv_pc, ori_pc, k=16)   -> [b,n]         Lib/loss_utils.py:99-107
    corresponding_normal_loss(adv_pc, normal, k=2) -> [b,n]    Lib/loss_utils.py:109-117
    repulsion_loss(pc, k=4, h=0.03)           -> [b,n]         Lib/loss_utils.py:119-123
    distance_kmean_loss(pc, k)                -> [b,n]         Lib/loss_utils.py:125-133
    kNN_smoothing_loss(adv_pc, k, threshold_coef=1.05) -> [b]  Lib/loss_utils.py:135-149
    uniform_loss(adv_pc, percentages, radius=1.0, k=2) -> []   Lib/loss_utils.py:151-190

All clouds are float32 CUDA tensors [b,3,n].  Where the reference launches four identical
adv->ori 1-NN searches per step this module runs the fused bidirectional kernel once
and shares the result through a per-step cache keyed on the adv tensor *object* (weakref + version
counter; never data_ptr, the caching allocator recycles addresses).  Gradients flow to adv_pc only
(ori_pc / normals / ori_kappa are constants of the attack, Attacker/geoA3_attack.py:196-217) and are
produced by one deterministic gather kernel per autograd node — no float atomics.

`geo_loss` is the fused fast path the attack driver uses: one autograd node for
w_cd*CD + w_hd*HD + w_curv*CUR (the composition of Attacker/geoA3_attack.py:131-162) with a single
backward launch.

The second group (displacement ... uniform) are the regularisers the reference keeps next to the
geometry-aware loss.  Their O(n^2) part — the neighbour search the reference does
with a dense [b,n,n] matrix + topk or with pytorch3d — runs on the exact top-K kernel (and, for
uniform_loss, on the FPS / ball_query / grouping kernels); the O(n*k) differentiable tail is the
reference's own tensor algebra on the gathered neighbours.

There is no CPU or pure-PyTorch fallback: a missing libgeoa3_b200.so or a CPU tensor raises.
"""
import math
import weakref

import torch

from . import ops

__all__ = ["norm_l2_loss", "chamfer_loss", "pseudo_chamfer_loss", "hausdorff_loss", "_get_kappa_ori",
           "_get_kappa_adv", "curvature_loss", "displacement_loss", "corresponding_normal_loss", "repulsion_loss",
           "distance_kmean_loss", "kNN_smoothing_loss", "uniform_loss", "geo_loss", "clear_cache", "HintBuffers", "NO_HINTS"]


# ---------------------------------------------------------------------------- search hints
class HintBuffers(object):
    """Persistent index buffers of one optimisation (owned by the attack driver).  The 1-NN / kNN kernels are
    exact for ANY seed, but a seed close to the answer makes them ~2-5x faster; across attack iterations
    the previous step's indices are such a seed.  The kernels read the hint and write the new result into
    the SAME buffer, so the hints stay fresh even when the whole step is replayed as a CUDA graph."""

    def __init__(self, prune_min_n_knn=2048, use_cells=True):
        self.d1 = self.jstar = self.d2 = self.istar = None
        self.nbr = {}
        # cell-grid searches (geoa3_cell_sort / geoa3_nn_pair_cells / geoa3_knn_cells): the original cloud's blob is
        # built once, the adversarial cloud's every step (one launch feeds the 1-NN and the kNN search).  Results are
        # identical to the scanning kernels; use_cells=False keeps those (A/B measurements, tests).
        self.use_cells = use_cells
        self.knn_k = 16  # neighbourhood size the adversarial grid is sized for (set by the fused loss node)
        self.cells_ori = self.cells_adv = None   # ops.Cells
        # visiting order for the pruned searches (ops.visit_order of the ORIGINAL cloud), computed once (first call)
        self.perm = self.iperm = self.ori_arranged = None
        self.prune_min_n_knn = prune_min_n_knn  # measured: box pruning of the kNN scan only pays for larger clouds
        # measurement hook (bench.py / tools): when set to a dict {"jstar","istar","nbr":{k:..}} of SAVED indices the
        # launches read their hints from it instead of from the buffers they refresh, so a launch can be repeated
        # on the same (cloud, previous-step hint) pair.  Results never depend on hints, so this changes timing only.
        self.frozen = None
        self._side = None

    def side_stream(self, device):
        """Stream the attack step runs the geometry losses on, next to the victim network (attack.forward_step)."""
        if self._side is None or self._side.device != torch.device(device):
            self._side = torch.cuda.Stream(device=device)
        return self._side

    def ensure_order(self, ori):
        if self.perm is None or self.perm.shape != (ori.shape[0], ori.shape[2]):
            self.perm, self.iperm = ops.visit_order(ori)
            self.ori_arranged = ops.arrange(ori, self.perm)

    def refresh_order(self, ori):
        """The original cloud's CONTENTS changed (AttackState.load_batch): recompute the visiting order into the
        existing tensors (their addresses are baked into a captured CUDA graph)."""
        self.refresh_cells(ori)
        if self.perm is None or self.perm.shape != (ori.shape[0], ori.shape[2]):
            return  # nothing computed yet; ensure_order() will do it on first use
        perm, iperm = ops.visit_order(ori)
        self.perm.copy_(perm)
        self.iperm.copy_(iperm)
        self.ori_arranged.copy_(ops.arrange(ori, self.perm))

    CELLS_MAX_N = 3072  # the kNN CTA stages blob + lists in shared memory with >= 2 CTAs per SM up to here

    def cells_ok(self, n, m):
        return self.use_cells and 32 <= n <= self.CELLS_MAX_N and 32 <= m <= self.CELLS_MAX_N

    KREF_NN = 4.0  # the 1-NN balls of an attack step are small: the original cloud gets a finer grid (measured)

    def ensure_cells(self, ori):
        """Cells of the ORIGINAL cloud (static during an attack)."""
        b, _, m = ori.shape
        if self.cells_ori is None or self.cells_ori.blobs.shape[0] != b or self.cells_ori.n != m:
            self.cells_ori = ops.cell_sort(ori, kref=self.KREF_NN)

    def refresh_cells(self, ori):
        c = self.cells_ori
        if c is not None and c.blobs.shape[0] == ori.shape[0] and c.n == ori.shape[2]:
            ops.cell_sort(ori, kref=self.KREF_NN, out=c)

    def sort_adv(self, adv, k=16):
        """Cells of the adversarial cloud, rebuilt every step into the same buffer (grid sized for the kNN balls)."""
        b, _, n = adv.shape
        c = self.cells_adv
        if c is None or c.blobs.shape[0] != b or c.n != n:
            self.cells_adv = ops.cell_sort(adv, kref=k + 1)
        else:
            ops.cell_sort(adv, kref=k + 1, out=c)
        return self.cells_adv

    def ensure_nn(self, b, n, m, dev):
        if self.jstar is None or self.jstar.shape != (b, n) or self.istar.shape != (b, m):
            self.d1 = torch.empty(b, n, device=dev, dtype=torch.float32)
            self.d2 = torch.empty(b, m, device=dev, dtype=torch.float32)
            self.jstar = torch.arange(n, device=dev, dtype=torch.int32).clamp_(max=m - 1).repeat(b, 1)
            self.istar = torch.arange(m, device=dev, dtype=torch.int32).clamp_(max=n - 1).repeat(b, 1)


NO_HINTS = object()  # pass as `hints` when consecutive calls see unrelated clouds (e.g. a fresh random subsample each
                     # step): stale indices are still exact as seeds, but they cost more than searching unseeded
_LAST = {}  # (device, b, n, m) -> last results, reused as (non-aliased) hints by the plain reference API
_ORDER = {}  # id(ori) -> (weakref(ori), version, perm, iperm, ori_arranged): visiting order of a cloud seen before


USE_CELLS = True  # plain reference API: cell-grid searches for clouds of 32..CELLS_MAX_N points (A/B switch for tests)
_CELLS = {}  # id(ori) -> (weakref(ori), version, Cells): cell grid of an original cloud seen before


def _cells_of(ori_obj, ori_c):
    ent = _CELLS.get(id(ori_obj))
    if ent is not None and ent[0]() is ori_obj and ent[1] == ori_obj._version:
        return ent[2]
    cells = ops.cell_sort(ori_c, kref=HintBuffers.KREF_NN)
    if len(_CELLS) > 8:
        _CELLS.clear()
    _CELLS[id(ori_obj)] = (weakref.ref(ori_obj), ori_obj._version, cells)
    return cells


def _order_of(ori_obj, ori_c):
    """Visiting order of an original cloud, cached per tensor object (the attack passes the same pc_ori every
    step).  Pure accelerator: the searches are exact for any order."""
    ent = _ORDER.get(id(ori_obj))
    if ent is not None and ent[0]() is ori_obj and ent[1] == ori_obj._version:
        return ent[2], ent[3], ent[4]
    perm, iperm = ops.visit_order(ori_c)
    arranged = ops.arrange(ori_c, perm)
    if len(_ORDER) > 8:
        _ORDER.clear()
    _ORDER[id(ori_obj)] = (weakref.ref(ori_obj), ori_obj._version, perm, iperm, arranged)
    return perm, iperm, arranged


# ---------------------------------------------------------------------------- per-step cache
class _Entry(object):
    __slots__ = ("adv_ref", "adv_ver", "ori_ref", "ori_ver", "adv_c", "ori_c", "d1", "jstar", "d2", "istar",
                 "red", "nbr", "kap", "hints", "arr", "cells")

    def matches(self, adv, ori):
        return (self.adv_ref() is adv and self.adv_ver == adv._version and self.ori_ref() is ori
                and self.ori_ver == ori._version)


# Cache contract: a hit needs the SAME tensor objects with unchanged `_version` counters.  In-place ops bump the
# version; swapping storage behind torch's back (`adv.data = other`, the reference's idiom at
# Attacker/geoA3_attack.py:345-352) does NOT — call clear_cache() after such a swap.  The caches are per process and
# not synchronised: use one attack per process / thread (the deployment model is one process per GPU).
_CACHE = []
_CACHE_MAX = 4


def clear_cache():
    del _CACHE[:]
    _LAST.clear()
    _ORDER.clear()
    _CELLS.clear()


def _as_input(t, name):
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor: the GeoA3 hot path has no CPU fallback" % name)
    if t.dim() != 3 or t.size(1) != 3:
        raise RuntimeError("%s must be [b,3,n]" % name)
    t = t.detach()
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _entry(adv, ori, hints=None):
    for e in _CACHE:
        if e.matches(adv, ori):
            return e
    e = _Entry()
    e.hints = hints
    e.adv_ref, e.adv_ver = weakref.ref(adv), adv._version
    e.ori_ref, e.ori_ver = weakref.ref(ori), ori._version
    e.adv_c, e.ori_c = _as_input(adv, "adv_pc"), _as_input(ori, "ori_pc")
    e.d1 = e.jstar = e.d2 = e.istar = e.red = e.arr = e.cells = None
    e.nbr, e.kap = {}, {}
    _CACHE.append(e)
    if len(_CACHE) > _CACHE_MAX:
        _CACHE.pop(0)
    return e


def _nn(e, both):
    """fused 1-NN search, computed once per (adv, ori) pair and step"""
    if e.d1 is None or (both and e.d2 is None):
        b, _, n = e.adv_c.shape
        m = e.ori_c.shape[2]
        hb = e.hints
        if hb is NO_HINTS:
            e.d1, e.jstar, e.d2, e.istar = ops.nn_pair(e.adv_c, e.ori_c, both=True)
        elif hb is not None:  # persistent buffers: hint and result alias, refreshed in place
            _launch_nn_hinted(e, hb)
        else:
            key = (e.adv_c.device, b, n, m)
            prev = _LAST.get(key)
            if USE_CELLS and 32 <= n <= HintBuffers.CELLS_MAX_N and 32 <= m <= HintBuffers.CELLS_MAX_N and e.ori_ref() is not None:
                # the reference API called step by step (chamfer_loss, hausdorff_loss, _get_kappa_adv ... on the same
                # clouds): the cell-grid searches, seeded with the previous call's results
                e.cells = ops.cell_sort(e.adv_c, kref=17.0)
                e.d1, e.jstar, e.d2, e.istar = ops.nn_pair_cells(e.cells, _cells_of(e.ori_ref(), e.ori_c),
                                                                 hint_a2o=prev[0] if prev else None,
                                                                 hint_o2a=prev[1] if prev else None)
            else:
                kw = {}
                if n == m and e.ori_ref() is not None:
                    perm, iperm, arranged = _order_of(e.ori_ref(), e.ori_c)
                    kw = dict(perm_a=perm, perm_o=perm, iperm_a=iperm, iperm_o=iperm, ori_arranged=arranged)
                e.d1, e.jstar, e.d2, e.istar = ops.nn_pair(e.adv_c, e.ori_c, both=True,
                                                           hint_a2o=prev[0] if prev else None,
                                                           hint_o2a=prev[1] if prev else None, **kw)
            _LAST[key] = (e.jstar, e.istar)
        e.red = None
    return e


def _reductions(e, one_sided=False):
    """CD / HD(+argmax) of the cached 1-NN distances; the one-sided mean (pseudo_chamfer_loss) only on demand."""
    if e.red is None:
        _nn(e, True)
        e.red = ops.kappa_loss_fwd(e.adv_c, d_a2o=e.d1, d_o2a=e.d2, m=e.ori_c.shape[2], want_kappa=False,
                                   want_cd=True, want_hd=True)
    if one_sided and "cd_one_sided" not in e.red:
        e.red["cd_one_sided"] = ops.kappa_loss_fwd(e.adv_c, d_a2o=e.d1, d_o2a=None, m=e.ori_c.shape[2],
                                                   want_kappa=False, want_cd=True)["cd"]
    return e.red


def _nbr(e, k):
    if k not in e.nbr:
        hb = e.hints
        if hb is NO_HINTS:
            e.nbr[k] = ops.knn(e.adv_c, e.adv_c, k + 1, drop=1)[0]
        elif hb is not None:
            buf = hb.nbr.get(k)
            if buf is None or buf.shape[:2] != e.adv_c.shape[::2]:
                # first call: nothing to hint with — the plain sorted search seeds the buffer, the regular launch
                # then rewrites it in its own (visiting) order, so step 0 sums kappa in the same order as every later step
                hb.nbr[k] = ops.knn(e.adv_c, e.adv_c, k + 1, drop=1)[0]
            _launch_knn_hinted(e, k, hb)
            e.nbr[k] = hb.nbr[k]
        else:
            key = (e.adv_c.device,) + tuple(e.adv_c.shape) + (k,)
            prev = _LAST.get(key)
            if e.cells is not None and prev is not None:   # member sets through this step's cell grid (kappa only sums)
                e.nbr[k] = ops.knn_cells(e.cells, k + 1, drop=1, hint=prev)[0]
            else:
                e.nbr[k] = ops.knn(e.adv_c, e.adv_c, k + 1, drop=1, hint=prev)[0]
            _LAST[key] = e.nbr[k]
    return e.nbr[k]


# ---------------------------------------------------------------------------- the launches of one attack step
# Each own kernel of the steady-state step is issued from exactly one function, used both by the autograd nodes
# below and by `step_plan` (bench.py / tools time these closures, so what is timed is what the attack runs).
def _launch_cell_sort(e, hb):
    """The adversarial cloud's cell grid of this step (geoa3_cell_sort), rebuilt into the persistent buffer."""
    hb.ensure_cells(e.ori_c)
    e.cells = hb.sort_adv(e.adv_c, hb.knn_k)


def _launch_nn_hinted(e, hb):
    b, _, n = e.adv_c.shape
    m = e.ori_c.shape[2]
    hb.ensure_nn(b, n, m, e.adv_c.device)
    hj, hi = (hb.frozen["jstar"], hb.frozen["istar"]) if hb.frozen else (hb.jstar, hb.istar)
    if hb.cells_ok(n, m):
        # cell-grid path: ONE launch sorts adv into its cell grid (also used by the kNN search of this step); each query
        # then only meets the candidates within reach of the distance to last step's argmin
        if getattr(e, "cells", None) is None:
            _launch_cell_sort(e, hb)
        e.arr = None
        ops.nn_pair_cells(e.cells, hb.cells_ori, hint_a2o=hj, hint_o2a=hi, out=(hb.d1, hb.jstar, hb.d2, hb.istar))
    elif n == m:  # both clouds share the visiting order of the original cloud (adv_i is a perturbed ori_i)
        hb.ensure_order(e.ori_c)
        # ONE launch arranges adv into the visiting order and boxes its groups: both pruned searches of the step use it
        e.arr = ops.arrange(e.adv_c, hb.perm, with_bbox=True)
        ops.nn_pair(e.adv_c, e.ori_c, hint_a2o=hj, hint_o2a=hi, perm_a=hb.perm, perm_o=hb.perm,
                    iperm_a=hb.iperm, iperm_o=hb.iperm, ori_arranged=hb.ori_arranged, adv_arranged=e.arr[0],
                    out=(hb.d1, hb.jstar, hb.d2, hb.istar))
    else:
        e.arr = None
        ops.nn_pair(e.adv_c, e.ori_c, hint_a2o=hj, hint_o2a=hi, out=(hb.d1, hb.jstar, hb.d2, hb.istar))
    e.d1, e.jstar, e.d2, e.istar = hb.d1, hb.jstar, hb.d2, hb.istar


def _launch_knn_hinted(e, k, hb):
    """Neighbour lists of the curvature term, refreshed in place (hint and output are the same buffer).  kappa and its
    gradient only sum over the neighbourhood, so clouds that fit one staged chunk use the member-set kernel
    (geoa3_knn_set: same members, visiting order) on the arrangement nn_pair already made this step; larger clouds
    keep the sorted scan, box-pruned where that was measured to pay (n >= 2048, K <= 17)."""
    buf = hb.nbr[k]
    hint = hb.frozen["nbr"][k] if hb.frozen else buf
    n = e.adv_c.shape[2]
    cells = getattr(e, "cells", None)
    if cells is not None:
        ops.knn_cells(cells, k + 1, drop=1, hint=hint, out=buf)
        return
    ordered = hb.perm is not None and n == hb.perm.shape[1]
    arr = getattr(e, "arr", None)
    if n <= 2048:
        if ordered and arr is not None:
            ops.knn(e.adv_c, e.adv_c, k + 1, drop=1, hint=hint, out=buf, perm_q=hb.perm, perm_c=hb.perm, iperm_c=hb.iperm,
                    arranged=arr, members_only=True)
        else:
            ops.knn(e.adv_c, e.adv_c, k + 1, drop=1, hint=hint, out=buf, members_only=True)
    elif ordered and k <= 16 and n >= hb.prune_min_n_knn:
        ops.knn(e.adv_c, e.adv_c, k + 1, drop=1, hint=hint, out=buf, perm_q=hb.perm, perm_c=hb.perm, iperm_c=hb.iperm,
                arranged=arr)
    else:
        ops.knn(e.adv_c, e.adv_c, k + 1, drop=1, hint=hint, out=buf)


def _launch_geo_fwd(e, nrm_src, kappa_ori, nbr, single_side, use_curv):
    return ops.kappa_loss_fwd(
        e.adv_c, normal=nrm_src if use_curv else None, jstar=e.jstar, nbr=nbr, d_a2o=e.d1,
        d_o2a=None if single_side else e.d2, kappa_ori=kappa_ori if use_curv else None, m=e.ori_c.shape[2],
        want_kappa=use_curv, want_nrm=use_curv, want_cd=True, want_hd=True, want_curv=use_curv)


FUSE_FWD_BWD = True  # geo_loss: one geoa3_geo_fwd_bwd launch instead of kappa_loss_fwd + loss_bwd (A/B switch for tests)


def _launch_geo_fused(e, nrm_src, kappa_ori, nbr, single_side, w_cd, w_hd, w_cu):
    return ops.geo_fwd_bwd(e.adv_c, e.ori_c, nrm_src, kappa_ori, e.jstar, None if single_side else e.istar, nbr, e.d1,
                           None if single_side else e.d2, w_cd, w_hd, w_cu)


def _launch_geo_bwd(e, out, nbr, kappa_ori, g_cd, g_hd, g_cu, single_side):
    return ops.loss_bwd(e.adv_c, ori=e.ori_c, nrm_adv=out["nrm"], kappa_adv=out["kappa"], kappa_ori=kappa_ori,
                        jstar=e.jstar, istar=None if single_side else e.istar, nbr=nbr, hd_arg=out["hd_arg"],
                        g_cd=g_cd, g_hd=g_hd, g_cu=g_cu)


def step_plan(adv, ori, ori_normal, ori_kappa, k, hints, w=(1.0, 0.1, 1.0)):
    """Runs the fused loss forward of one attack step on `adv` through the persistent `hints` (exactly what
    `geo_loss(..., hints=hints)` launches) and returns {"out": forward results, "launches": f} where f(g) lists
    (name, closure) for every own kernel of the step's forward AND backward with upstream gradient g [b] —
    measurement hooks for bench.py and tools/; results are identical to the autograd path."""
    e = _Entry()
    e.hints = hints
    e.adv_ref = e.ori_ref = lambda: None
    e.adv_ver = e.ori_ver = -1
    e.adv_c, e.ori_c = _as_input(adv, "adv_pc"), _as_input(ori, "ori_pc")
    e.d1 = e.jstar = e.d2 = e.istar = e.red = e.arr = e.cells = None
    e.nbr, e.kap = {}, {}
    hints.knn_k = k
    nrm_src = _as_input(ori_normal, "ori_normal")
    ko = ori_kappa.detach().float().contiguous()
    _launch_nn_hinted(e, hints)
    nbr = _nbr(e, k)
    out = _launch_geo_fwd(e, nrm_src, ko, nbr, False, True)

    def launches(g):
        g = g.detach().float().contiguous()
        gs = [(g * w_) .contiguous() for w_ in w]
        pre = [("cell_sort", lambda: _launch_cell_sort(e, hints))] if e.cells is not None else []
        pre += [("nn_pair", lambda: _launch_nn_hinted(e, hints)), ("knn", lambda: _launch_knn_hinted(e, k, hints))]
        if FUSE_FWD_BWD and ops.geo_fwd_bwd_supported(e.adv_c.shape[2], e.ori_c.shape[2], k):
            return pre + [("geo_fwd_bwd", lambda: _launch_geo_fused(e, nrm_src, ko, nbr, False, w[0], w[1], w[2]))]
        return pre + [("kappa_loss_fwd", lambda: _launch_geo_fwd(e, nrm_src, ko, nbr, False, True)),
                      ("loss_bwd", lambda: _launch_geo_bwd(e, out, nbr, ko, gs[0], gs[1], gs[2], False))]

    return {"out": out, "entry": e, "nbr": nbr, "launches": launches}



def _grad_vec(g, like):
    return g.detach().to(torch.float32).contiguous()


# ---------------------------------------------------------------------------- autograd nodes
class _Chamfer(torch.autograd.Function):
    @staticmethod
    def forward(ctx, adv_pc, ori_pc, both):
        e = _nn(_entry(adv_pc, ori_pc), True)
        red = _reductions(e, one_sided=not both)
        ctx.e, ctx.both = e, both
        return (red["cd"] if both else red["cd_one_sided"]).clone()

    @staticmethod
    def backward(ctx, g):
        e = ctx.e
        grad = ops.loss_bwd(e.adv_c, ori=e.ori_c, jstar=e.jstar, istar=e.istar if ctx.both else None,
                            g_cd=_grad_vec(g, e.adv_c))
        return grad, None, None


class _Hausdorff(torch.autograd.Function):
    @staticmethod
    def forward(ctx, adv_pc, ori_pc):
        e = _nn(_entry(adv_pc, ori_pc), True)
        red = _reductions(e)
        ctx.e = e
        return red["hd"].clone()

    @staticmethod
    def backward(ctx, g):
        e = ctx.e
        grad = ops.loss_bwd(e.adv_c, ori=e.ori_c, jstar=e.jstar, hd_arg=e.red["hd_arg"], g_hd=_grad_vec(g, e.adv_c))
        return grad, None


class _KappaAdv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, adv_pc, ori_pc, ori_normal, k):
        e = _nn(_entry(adv_pc, ori_pc), True)
        nbr = _nbr(e, k)
        nrm_src = _as_input(ori_normal, "ori_normal")
        out = ops.kappa_loss_fwd(e.adv_c, normal=nrm_src, jstar=e.jstar, nbr=nbr, want_kappa=True, want_nrm=True)
        ctx.e, ctx.nbr, ctx.nrm = e, nbr, out["nrm"]
        ctx.mark_non_differentiable(out["nrm"])
        return out["kappa"], out["nrm"]

    @staticmethod
    def backward(ctx, g_kappa, _g_nrm):
        e = ctx.e
        grad = ops.loss_bwd(e.adv_c, nrm_adv=ctx.nrm, nbr=ctx.nbr, g_kappa=_grad_vec(g_kappa, e.adv_c),
                            m=e.ori_c.shape[2])
        return grad, None, None, None


class _KappaOri(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pc, normal, k):
        pc_c, nrm_c = _as_input(pc, "pc"), _as_input(normal, "normal")
        nbr = ops.knn(pc_c, pc_c, k + 1, drop=1)[0]
        out = ops.kappa_loss_fwd(pc_c, normal=nrm_c, jstar=None, nbr=nbr, want_kappa=True)
        ctx.pc, ctx.nrm, ctx.nbr = pc_c, nrm_c, nbr
        return out["kappa"]

    @staticmethod
    def backward(ctx, g_kappa):
        grad = ops.loss_bwd(ctx.pc, nrm_adv=ctx.nrm, nbr=ctx.nbr, g_kappa=_grad_vec(g_kappa, ctx.pc))
        return grad, None, None


class _GeoLoss(torch.autograd.Function):
    """w_cd*CD + w_hd*HD + w_curv*CUR in one node (Attacker/geoA3_attack.py:131-162)."""

    @staticmethod
    def forward(ctx, adv_pc, ori_pc, ori_normal, ori_kappa, k, w_cd, w_hd, w_curv, single_side, hints):
        if isinstance(hints, HintBuffers) and k > 0:
            hints.knn_k = k
        e = _nn(_entry(adv_pc, ori_pc, hints), True)
        use_curv = w_curv != 0 and k > 0
        nbr = _nbr(e, k) if use_curv else None
        ko = ori_kappa.detach().float().contiguous() if use_curv else None
        n, m = e.adv_c.shape[2], e.ori_c.shape[2]
        ctx.fused = bool(ctx.needs_input_grad[0]) and FUSE_FWD_BWD and ops.geo_fwd_bwd_supported(n, m, k if use_curv else 0)
        if ctx.fused:
            # forward and backward in ONE launch: the gradient for a unit upstream gradient is computed now,
            # backward() only scales it (the loss is linear in its upstream gradient)
            out = _launch_geo_fused(e, _as_input(ori_normal, "ori_normal") if use_curv else None, ko, nbr, single_side,
                                    w_cd, w_hd, w_curv if use_curv else 0.0)
            cd, hd, curv = out["cd"], out["hd"], out["curv"]
            ctx.unit_grad = out["grad"]
        else:
            out = _launch_geo_fwd(e, _as_input(ori_normal, "ori_normal") if use_curv else None, ko, nbr, single_side, use_curv)
            cd, hd = out["cd"], out["hd"]
            curv = out["curv"] if use_curv else torch.zeros_like(cd)
            ctx.e, ctx.out, ctx.nbr = e, out, nbr
            ctx.kappa_ori = ko
        total = w_cd * cd + w_hd * hd + (w_curv * curv if use_curv else 0.0)
        ctx.w = (w_cd, w_hd, w_curv, single_side, use_curv)
        ctx.mark_non_differentiable(cd, hd, curv)
        return total, cd, hd, curv

    @staticmethod
    def backward(ctx, g, _a, _b, _c):
        if ctx.fused:
            ug = ctx.unit_grad
            g = g.detach().float().reshape(-1, 1, 1)
            return (ug * g,) + (None,) * 9
        e, out = ctx.e, ctx.out
        w_cd, w_hd, w_curv, single_side, use_curv = ctx.w
        g = _grad_vec(g, e.adv_c)
        grad = _launch_geo_bwd(e, out, ctx.nbr, ctx.kappa_ori, (g * w_cd) if w_cd != 0 else None,
                               (g * w_hd) if w_hd != 0 else None, (g * w_curv) if use_curv else None, single_side)
        return (grad,) + (None,) * 9


# ---------------------------------------------------------------------------- reference API
def norm_l2_loss(adv_pc, ori_pc):
    return ((adv_pc - ori_pc) ** 2).sum(1).sum(1)


def chamfer_loss(adv_pc, ori_pc):
    return _Chamfer.apply(adv_pc, ori_pc, True)


def pseudo_chamfer_loss(adv_pc, ori_pc):
    return _Chamfer.apply(adv_pc, ori_pc, False)


def hausdorff_loss(adv_pc, ori_pc):
    return _Hausdorff.apply(adv_pc, ori_pc)


def _get_kappa_ori(pc, normal, k=2):
    return _KappaOri.apply(pc, normal, k)


def _get_kappa_adv(adv_pc, ori_pc, ori_normal, k=2):
    return _KappaAdv.apply(adv_pc, ori_pc, ori_normal, k)


def curvature_loss(adv_pc, ori_pc, adv_kappa, ori_kappa, k=2):
    """mean_i (kappa_adv_i - kappa_ori[j*(i)])^2; `k` is unused, as in the reference (:84)."""
    e = _nn(_entry(adv_pc, ori_pc), True)
    onenn_ori_kappa = torch.gather(ori_kappa, 1, e.jstar.long())
    return ((adv_kappa - onenn_ori_kappa) ** 2).mean(-1)


# ---------------------------------------------------------------------------- neighbourhood regularisers
def _self_nbr(pc, k):
    """Self kNN (K = k+1, own entry dropped like the reference's `[:, :, 1:]`) -> int64 [b,n,k]; no gradient,
    exactly as topk / knn_points indices carry none."""
    c = _as_input(pc.detach(), "pc")
    return ops.knn(c, c, int(k) + 1, drop=1)[0].long()


def _nbr_vectors(pc, nbr):
    """pc [b,3,n], nbr [b,n,k] -> pc[nbr] - pc as [b,3,n,k], differentiable w.r.t. pc."""
    b, _, n = pc.shape
    k = nbr.shape[2]
    nn_pts = torch.gather(pc, 2, nbr.reshape(b, 1, n * k).expand(b, 3, n * k)).view(b, 3, n, k)
    return nn_pts - pc.unsqueeze(3)


def displacement_loss(adv_pc, ori_pc, k=16):
    """Variance-like penalty on the displacement magnitude among the k nearest ORIGINAL neighbours."""
    b, _, n = adv_pc.shape
    inter_idx = _self_nbr(ori_pc, k)
    theta_distance = ((adv_pc - ori_pc) ** 2).sum(1)
    nn_theta = torch.gather(theta_distance, 1, inter_idx.view(b, n * k)).view(b, n, k)
    return ((nn_theta - theta_distance.unsqueeze(2)) ** 2).mean(2)


def corresponding_normal_loss(adv_pc, normal, k=2):
    """mean_m |<normal_i, unit(adv[nbr(i,m)] - adv_i)>| with the point's own normal."""
    from .utility import _normalize

    vectors = _normalize(_nbr_vectors(adv_pc, _self_nbr(adv_pc, k)))
    return torch.abs((vectors * normal.unsqueeze(3)).sum(1)).mean(2)


def repulsion_loss(pc, k=4, h=0.03):
    dis = (_nbr_vectors(pc, _self_nbr(pc, k)) ** 2).sum(1)
    return -(dis * torch.exp(-(dis ** 2) / (h ** 2))).mean(2)


def distance_kmean_loss(pc, k):
    b, _, n = pc.shape
    nbr = _self_nbr(pc, k)
    # the reference measures |p_i - p_j + 1e-12| (the epsilon is added to every coordinate difference, :127)
    dis = ((1e-12 - _nbr_vectors(pc, nbr)) ** 2).sum(1).sqrt()
    dis_mean = dis.mean(-1)
    dis_mean_k = torch.gather(dis_mean, 1, nbr.view(b, n * k)).view(b, n, k)
    return torch.abs(dis_mean.unsqueeze(2) - dis_mean_k).mean(-1)


def kNN_smoothing_loss(adv_pc, k, threshold_coef=1.05):
    knn_dis = (_nbr_vectors(adv_pc, _self_nbr(adv_pc, k)) ** 2).sum(1).mean(-1)
    threshold = knn_dis.mean(-1) + threshold_coef * knn_dis.std(-1)
    condition = torch.gt(knn_dis, threshold.unsqueeze(1)).float()
    return (knn_dis * condition).mean(1)


def uniform_loss(adv_pc, percentages=[0.004, 0.006, 0.008, 0.010, 0.012], radius=1.0, k=2):
    """PU-GAN style uniformity term.  The reference body (Lib/loss_utils.py:151-190) needs `pointnet2_utils`,
    which that file never imports; this is the function it evidently intends, on this package's own
    pointnet2_ops.  The FPS seeds do not depend on the percentage and are computed once."""
    from .pointnet2_ops import pointnet2_utils

    if adv_pc.size(1) == 3:
        adv_pc = adv_pc.permute(0, 2, 1).contiguous()
    b, n, _ = adv_pc.size()
    npoint = int(n * 0.05)
    adv_pc_flipped = adv_pc.transpose(1, 2).contiguous()
    seeds = pointnet2_utils.furthest_point_sample(adv_pc, npoint)
    new_xyz = pointnet2_utils.gather_operation(adv_pc_flipped, seeds).transpose(1, 2).contiguous()
    loss = None
    for p in percentages:
        p = p * 4
        nsample = int(n * p)
        r = math.sqrt(p * radius)
        disk_area = math.pi * (radius ** 2) * p / nsample
        expect_len = float(torch.sqrt(torch.tensor(disk_area, dtype=torch.float32)))  # fp32 like :161, host-side
        idx = pointnet2_utils.ball_query(r, nsample, adv_pc, new_xyz)
        grouped = pointnet2_utils.grouping_operation(adv_pc_flipped, idx)          # [b,3,npoint,nsample]
        grouped = grouped.permute(0, 2, 1, 3).reshape(b * npoint, 3, nsample)     # one small cloud per patch
        uniform_dis = (_nbr_vectors(grouped, _self_nbr(grouped, k)) ** 2).sum(1)    # [b*npoint,nsample,k]
        uniform_dis = torch.sqrt(torch.abs(uniform_dis) + 1e-12).mean(-1)
        uniform_dis = (uniform_dis - expect_len) ** 2 / (expect_len + 1e-12)
        mean = uniform_dis.mean() * math.pow(p * 100, 2)
        loss = mean if loss is None else loss + mean
    return loss / len(percentages)


def geo_loss(adv_pc, ori_pc, ori_normal, ori_kappa, k=16, w_cd=1.0, w_hd=0.1, w_curv=1.0, single_side=False,
             hints=None):
    """Fused constrain loss. Returns (w_cd*CD + w_hd*HD + w_curv*CUR [b], CD [b], HD [b], CUR [b]);
    only the first output carries gradient.  `hints` (HintBuffers, optional) carries the previous step's
    indices as search seeds — a pure accelerator, results do not depend on it."""
    return _GeoLoss.apply(adv_pc, ori_pc, ori_normal, ori_kappa, int(k), float(w_cd), float(w_hd), float(w_curv),
                          bool(single_side), hints)
