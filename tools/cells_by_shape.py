"""Per-shape cost of the cell-grid searches (all 250 clouds of ONE synthetic family, hints = kNN of a cloud one
0.01-step away): shows which cloud shapes are expensive.   python tools/cells_by_shape.py [--grid 12]"""
import argparse
import json
import os.path as osp
import sys

import numpy as np
import torch

ROOT = osp.dirname(osp.dirname(osp.abspath(__file__)))
sys.path.insert(0, ROOT)
from geoa3_b200 import ops, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--kref", type=float, default=17)
ap.add_argument("--kori", type=float, default=4)
a = ap.parse_args()
b, n, k = 250, 1024, 16
nf = len(synth._FAMILIES)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def t(fn, iters=5):
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return round(sorted(ts)[len(ts) // 2], 1)


for fam in list(range(nf)) + [-1]:
    ids = [fam + nf * i for i in range(b)] if fam >= 0 else list(range(b))
    pcs = np.stack([synth.make_instance(i, n)[0] for i in ids[:50]])
    ori = torch.from_numpy(np.tile(pcs, (5, 1, 1))).cuda()
    g = torch.Generator(device="cuda").manual_seed(1)
    prev = ori + 0.03 * torch.randn(ori.shape, device="cuda", generator=g)
    adv = (prev + 0.01 * torch.sign(torch.randn(ori.shape, device="cuda", generator=g))).contiguous()
    hint = ops.knn(prev, prev, k + 1, drop=1)[0]
    j0, i0 = ops.nn_pair(prev, ori)[1::2]
    ba, bo = ops.cell_sort(adv, kref=a.kref), ops.cell_sort(ori, kref=a.kori)
    out = torch.empty_like(hint)
    grid = ba.blobs[0, 52:64].view(torch.float32).tolist()
    r = dict(family=fam, grid_of_cloud0=grid, sort=t(lambda: ops.cell_sort(adv, kref=a.kref, out=ba)),
             nn=t(lambda: ops.nn_pair_cells(ba, bo, hint_a2o=j0, hint_o2a=i0)),
             knn=t(lambda: ops.knn_cells(ba, k + 1, drop=1, hint=hint, out=out)))
    print(json.dumps(r), flush=True)
