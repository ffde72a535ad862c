"""bench.py's `loss_fwd_bwd_us` record alone (own kernels of one steady-state step on a REAL attack state, CUDA events,
L2 flushed before every launch) — for A/B builds (GEOA3_SO_PATH).  The state comes from a saved file when --load is
given (tools/cells_states.py --save), so a variant library is never used to produce it.
    python tools/loss_breakdown.py [--load /tmp/states.pt]"""
import argparse
import json
import os.path as osp
import sys

import torch

ROOT = osp.dirname(osp.dirname(osp.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from geoa3_b200 import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--load", type=str, default=None)
a = ap.parse_args()
dev = torch.device("cuda", 0)
if a.load:
    st = torch.load(a.load)
    ori = st[-1]["ori"].to(dev)
    prev, adv = st[-2]["adv"].to(dev), st[-1]["adv"].to(dev)
    import numpy as np
    nrm = torch.from_numpy(synth.make_batch(ori.shape[0], ori.shape[2], 0)[1]).to(dev)
else:
    s, pins = bench.build_state("PointNet", 250, bench.NPTS, 0, 250, dev)
    for _ in range(60):
        s.step()
    prev = (s.base + s.offset).detach().clone()
    s.step()
    adv = (s.base + s.offset).detach().clone()
    ori, nrm = s.pc_ori.detach(), s.normal_ori.detach()
kb = bench.kernel_breakdown(ori, nrm, prev, adv, bench.KNN)
kb["total"] = round(sum(kb.values()), 2)
print(json.dumps(kb))
