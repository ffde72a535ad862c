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

All clouds are float32 CUDA tensors [b,3,n].  Where the reference launches four identical
adv->ori 1-NN searches per step this module runs the fused bidirectional kernel once
and shares the result through a per-step cache keyed on the adv tensor *object* (weakref + version
counter; never data_ptr, the caching allocator recycles addresses).  Gradients flow to adv_pc only
(ori_pc / normals / ori_kappa are constants of the attack, Attacker/geoA3_attack.py:196-217) and are
produced by one deterministic gather kernel per autograd node — no float atomics.

`geo_loss` is the fused fast path the attack driver uses: one autograd node for
w_cd*CD + w_hd*HD + w_curv*CUR (the composition of Attacker/geoA3_attack.py:131-162) with a single
backward launch.

There is no CPU or pure-PyTorch fallback: a missing libgeoa3_b200.so or a CPU tensor raises.
"""
import weakref

import torch

from . import ops

__all__ = ["norm_l2_loss", "chamfer_loss", "pseudo_chamfer_loss", "hausdorff_loss", "_get_kappa_ori",
           "_get_kappa_adv", "curvature_loss", "geo_loss", "clear_cache", "HintBuffers"]


# ---------------------------------------------------------------------------- search hints
class HintBuffers(object):
    """Persistent index buffers of one optimisation (owned by the attack driver).  The 1-NN / kNN kernels are
    exact for ANY seed, but a seed close to the answer makes them ~2-5x faster; across attack iterations
    the previous step's indices are such a seed.  The kernels read the hint and write the new result into
    the SAME buffer, so the hints stay fresh even when the whole step is replayed as a CUDA graph."""

    def __init__(self, prune_min_n_knn=2048):
        self.d1 = self.jstar = self.d2 = self.istar = None
        self.nbr = {}
        # visiting order for the pruned searches: Morton order of the ORIGINAL cloud, computed once (first call)
        self.perm = self.iperm = self.ori_arranged = None
        self.prune_min_n_knn = prune_min_n_knn  # measured: box pruning of the kNN scan only pays for larger clouds

    def ensure_order(self, ori):
        if self.perm is None or self.perm.shape != (ori.shape[0], ori.shape[2]):
            self.perm, self.iperm = ops.morton_order(ori)
            self.ori_arranged = ops.arrange(ori, self.perm)

    def ensure_nn(self, b, n, m, dev):
        if self.jstar is None or self.jstar.shape != (b, n) or self.istar.shape != (b, m):
            self.d1 = torch.empty(b, n, device=dev, dtype=torch.float32)
            self.d2 = torch.empty(b, m, device=dev, dtype=torch.float32)
            self.jstar = torch.arange(n, device=dev, dtype=torch.int32).clamp_(max=m - 1).repeat(b, 1)
            self.istar = torch.arange(m, device=dev, dtype=torch.int32).clamp_(max=n - 1).repeat(b, 1)


_LAST = {}  # (device, b, n, m) -> last results, reused as (non-aliased) hints by the plain reference API
_ORDER = {}  # id(ori) -> (weakref(ori), version, perm, iperm, ori_arranged): Morton order of a cloud seen before


def _order_of(ori_obj, ori_c):
    """Visiting order of an original cloud, cached per tensor object (the attack passes the same pc_ori every
    step).  Pure accelerator: the searches are exact for any order."""
    ent = _ORDER.get(id(ori_obj))
    if ent is not None and ent[0]() is ori_obj and ent[1] == ori_obj._version:
        return ent[2], ent[3], ent[4]
    perm, iperm = ops.morton_order(ori_c)
    arranged = ops.arrange(ori_c, perm)
    if len(_ORDER) > 8:
        _ORDER.clear()
    _ORDER[id(ori_obj)] = (weakref.ref(ori_obj), ori_obj._version, perm, iperm, arranged)
    return perm, iperm, arranged


# ---------------------------------------------------------------------------- per-step cache
class _Entry(object):
    __slots__ = ("adv_ref", "adv_ver", "ori_ref", "ori_ver", "adv_c", "ori_c", "d1", "jstar", "d2", "istar",
                 "red", "nbr", "kap", "hints")

    def matches(self, adv, ori):
        return (self.adv_ref() is adv and self.adv_ver == adv._version and self.ori_ref() is ori
                and self.ori_ver == ori._version)


_CACHE = []
_CACHE_MAX = 4


def clear_cache():
    del _CACHE[:]
    _LAST.clear()
    _ORDER.clear()


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
    e.d1 = e.jstar = e.d2 = e.istar = e.red = None
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
        if hb is not None:  # persistent buffers: hint and result alias, refreshed in place
            hb.ensure_nn(b, n, m, e.adv_c.device)
            if n == m:  # both clouds share the visiting order of the original cloud (adv_i is a perturbed ori_i)
                hb.ensure_order(e.ori_c)
                ops.nn_pair(e.adv_c, e.ori_c, hint_a2o=hb.jstar, hint_o2a=hb.istar, perm_a=hb.perm, perm_o=hb.perm,
                            iperm_a=hb.iperm, iperm_o=hb.iperm, ori_arranged=hb.ori_arranged,
                            out=(hb.d1, hb.jstar, hb.d2, hb.istar))
            else:
                ops.nn_pair(e.adv_c, e.ori_c, hint_a2o=hb.jstar, hint_o2a=hb.istar,
                            out=(hb.d1, hb.jstar, hb.d2, hb.istar))
            e.d1, e.jstar, e.d2, e.istar = hb.d1, hb.jstar, hb.d2, hb.istar
        else:
            key = (e.adv_c.device, b, n, m)
            prev = _LAST.get(key)
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


def _reductions(e):
    if e.red is None:
        _nn(e, True)
        e.red = ops.kappa_loss_fwd(e.adv_c, d_a2o=e.d1, d_o2a=e.d2, m=e.ori_c.shape[2], want_kappa=False,
                                   want_cd=True, want_hd=True)
        one = ops.kappa_loss_fwd(e.adv_c, d_a2o=e.d1, d_o2a=None, m=e.ori_c.shape[2], want_kappa=False, want_cd=True)
        e.red["cd_one_sided"] = one["cd"]
    return e.red


def _nbr(e, k):
    if k not in e.nbr:
        hb = e.hints
        if hb is not None:
            buf = hb.nbr.get(k)
            if buf is None or buf.shape[:2] != e.adv_c.shape[::2]:
                hb.nbr[k] = ops.knn(e.adv_c, e.adv_c, k + 1, drop=1)[0]  # first call: nothing to hint with
            elif (hb.perm is not None and k <= 16 and e.adv_c.shape[2] >= hb.prune_min_n_knn
                  and e.adv_c.shape[2] == hb.perm.shape[1]):  # measured: pays for n >= 2048 and K <= 17 only
                ops.knn(e.adv_c, e.adv_c, k + 1, drop=1, hint=buf, out=buf, perm_q=hb.perm, perm_c=hb.perm,
                        iperm_c=hb.iperm)
            else:
                ops.knn(e.adv_c, e.adv_c, k + 1, drop=1, hint=buf, out=buf)
            e.nbr[k] = hb.nbr[k]
        else:
            key = (e.adv_c.device,) + tuple(e.adv_c.shape) + (k,)
            e.nbr[k] = ops.knn(e.adv_c, e.adv_c, k + 1, drop=1, hint=_LAST.get(key))[0]
            _LAST[key] = e.nbr[k]
    return e.nbr[k]


def _grad_vec(g, like):
    return g.detach().to(torch.float32).contiguous()


# ---------------------------------------------------------------------------- autograd nodes
class _Chamfer(torch.autograd.Function):
    @staticmethod
    def forward(ctx, adv_pc, ori_pc, both):
        e = _nn(_entry(adv_pc, ori_pc), True)
        red = _reductions(e)
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
        e = _nn(_entry(adv_pc, ori_pc, hints), True)
        use_curv = w_curv != 0 and k > 0
        nbr = _nbr(e, k) if use_curv else None
        out = ops.kappa_loss_fwd(
            e.adv_c, normal=_as_input(ori_normal, "ori_normal") if use_curv else None, jstar=e.jstar, nbr=nbr,
            d_a2o=e.d1, d_o2a=None if single_side else e.d2,
            kappa_ori=ori_kappa.detach().float().contiguous() if use_curv else None, m=e.ori_c.shape[2],
            want_kappa=use_curv, want_nrm=use_curv, want_cd=True, want_hd=True, want_curv=use_curv)
        cd, hd = out["cd"], out["hd"]
        curv = out["curv"] if use_curv else torch.zeros_like(cd)
        total = w_cd * cd + w_hd * hd + (w_curv * curv if use_curv else 0.0)
        ctx.e, ctx.out, ctx.nbr = e, out, nbr
        ctx.w = (w_cd, w_hd, w_curv, single_side, use_curv)
        ctx.kappa_ori = ori_kappa.detach().float().contiguous() if use_curv else None
        ctx.mark_non_differentiable(cd, hd, curv)
        return total, cd, hd, curv

    @staticmethod
    def backward(ctx, g, _a, _b, _c):
        e, out = ctx.e, ctx.out
        w_cd, w_hd, w_curv, single_side, use_curv = ctx.w
        g = _grad_vec(g, e.adv_c)
        grad = ops.loss_bwd(
            e.adv_c, ori=e.ori_c, nrm_adv=out["nrm"], kappa_adv=out["kappa"], kappa_ori=ctx.kappa_ori, jstar=e.jstar,
            istar=None if single_side else e.istar, nbr=ctx.nbr, hd_arg=out["hd_arg"],
            g_cd=(g * w_cd) if w_cd != 0 else None, g_hd=(g * w_hd) if w_hd != 0 else None,
            g_cu=(g * w_curv) if use_curv else None)
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


def geo_loss(adv_pc, ori_pc, ori_normal, ori_kappa, k=16, w_cd=1.0, w_hd=0.1, w_curv=1.0, single_side=False,
             hints=None):
    """Fused constrain loss. Returns (w_cd*CD + w_hd*HD + w_curv*CUR [b], CD [b], HD [b], CUR [b]);
    only the first output carries gradient.  `hints` (HintBuffers, optional) carries the previous step's
    indices as search seeds — a pure accelerator, results do not depend on it."""
    return _GeoLoss.apply(adv_pc, ori_pc, ori_normal, ori_kappa, int(k), float(w_cd), float(w_hd), float(w_curv),
                          bool(single_side), hints)
