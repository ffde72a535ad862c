"""Synthetic ModelNet-shaped clouds with analytic unit normals — the data side of the hot path
when no dataset is on disk (bench, smoke, tests; `provider.write_synthetic_mat` stores them in the
reference's .mat layout).

N surface samples of an analytic shape, centred and scaled to max-norm 1 like pc_normalize
(/root/reference/Provider/modelnet_trn_test.py:13-19); 10 shape families <-> the 10 class ids of
/root/reference/Provider/modelnet10_instance250.py:10.  Instance i uses a CPU generator seeded
20260+i.  Pure numpy, deterministic, no reference code involved.
"""
import numpy as np

CLASS_IDS = [17, 9, 36, 20, 3, 16, 34, 38, 23, 15]


def _unit(v):
    return v / np.maximum(np.linalg.norm(v, axis=-1, keepdims=True), 1e-12)


def _sphere(rng, n, ax=(1.0, 1.0, 1.0)):
    u = _unit(rng.standard_normal((n, 3)))
    ax = np.asarray(ax)
    return u * ax, _unit(u / ax)


def _box(rng, n, h=(1.0, 0.7, 0.4)):
    h = np.asarray(h)
    area = np.array([h[1] * h[2], h[0] * h[2], h[0] * h[1]])
    face = rng.choice(3, size=n, p=area / area.sum())
    sign = rng.choice([-1.0, 1.0], size=n)
    p = rng.uniform(-1, 1, (n, 3)) * h
    nr = np.zeros((n, 3))
    p[np.arange(n), face] = sign * h[face]
    nr[np.arange(n), face] = sign
    return p, nr


def _cylinder(rng, n, r=0.5, hh=1.0):
    t = rng.uniform(0, 2 * np.pi, n)
    side = rng.uniform(size=n) < (2 * hh) / (2 * hh + r)
    z = rng.uniform(-hh, hh, n)
    rr = np.where(side, r, r * np.sqrt(rng.uniform(size=n)))
    top = rng.choice([-1.0, 1.0], size=n)
    p = np.stack([rr * np.cos(t), rr * np.sin(t), np.where(side, z, top * hh)], 1)
    nr = np.where(side[:, None], np.stack([np.cos(t), np.sin(t), 0 * t], 1),
                  np.stack([0 * t, 0 * t, top], 1))
    return p, nr


def _torus(rng, n, R=1.0, r=0.35):
    u, v = rng.uniform(0, 2 * np.pi, n), rng.uniform(0, 2 * np.pi, n)
    p = np.stack([(R + r * np.cos(v)) * np.cos(u), (R + r * np.cos(v)) * np.sin(u), r * np.sin(v)], 1)
    nr = np.stack([np.cos(v) * np.cos(u), np.cos(v) * np.sin(u), np.sin(v)], 1)
    return p, nr


def _cone(rng, n, r=0.7, h=1.4):
    t = rng.uniform(0, 2 * np.pi, n)
    s = np.sqrt(rng.uniform(size=n))
    p = np.stack([r * s * np.cos(t), r * s * np.sin(t), h * (1 - s) - h / 2], 1)
    nr = _unit(np.stack([h * np.cos(t), h * np.sin(t), r + 0 * t], 1))
    return p, nr


def _plane(rng, n):  # flat plate through the centroid: exercises the FPS |p|^2<=1e-3 skip
    p = np.concatenate([rng.uniform(-1, 1, (n, 2)), np.zeros((n, 1))], 1)
    p[: max(1, n // 64)] *= 0.02
    nr = np.tile(np.array([[0.0, 0.0, 1.0]]), (n, 1))
    return p, nr


def _two_spheres(rng, n):
    p1, n1 = _sphere(rng, n // 2)
    p2, n2 = _sphere(rng, n - n // 2, (0.5, 0.5, 0.5))
    return np.concatenate([p1 + [0.8, 0, 0], p2 - [1.0, 0, 0]]), np.concatenate([n1, n2])


_FAMILIES = [
    lambda g, n: _sphere(g, n),
    lambda g, n: _sphere(g, n, (1.0, 0.6, 0.35)),
    lambda g, n: _box(g, n),
    lambda g, n: _cylinder(g, n),
    lambda g, n: _torus(g, n),
    lambda g, n: _cone(g, n),
    lambda g, n: _plane(g, n),
    lambda g, n: _two_spheres(g, n),
    lambda g, n: _box(g, n, (1.0, 1.0, 0.1)),
    lambda g, n: _cylinder(g, n, 0.15, 1.0),
]


def make_instance(i, n=1024):
    """-> (pc [3,n] f32, normal [3,n] f32, class id)"""
    rng = np.random.default_rng(20260 + i)
    fam = i % len(_FAMILIES)
    p, nr = _FAMILIES[fam](rng, n)
    perm = rng.permutation(n)
    p, nr = p[perm], nr[perm]
    p = p - p.mean(0, keepdims=True)
    p = p / np.sqrt((p ** 2).sum(1)).max()
    return p.T.astype(np.float32).copy(), _unit(nr).T.astype(np.float32).copy(), CLASS_IDS[fam]


def make_batch(b, n=1024, start=0):
    pcs, nrs, lab = zip(*[make_instance(start + i, n) for i in range(b)])
    return np.stack(pcs), np.stack(nrs), np.asarray(lab, np.int64)


def make_offsets(b, n, seed=0, std=1e-3):
    """N(0, 1e-3) initial perturbation (/root/reference/Attacker/geoA3_attack.py:265-267)."""
    return (np.random.default_rng(seed).standard_normal((b, 3, n)) * std).astype(np.float32)


def lattice_cloud(n=216):
    """Adversarial-tie fixture: points on an integer lattice => many exactly equal distances."""
    s = int(round(n ** (1 / 3)))
    g = np.stack(np.meshgrid(*[np.arange(s)] * 3, indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    g = (g - g.mean(0)) / 4.0
    nr = _unit(g + 1e-3)
    return g.T.copy(), nr.T.astype(np.float32).copy()
