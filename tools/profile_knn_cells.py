"""ncu target for the cell-grid kNN on a REAL attack state: runs `--steps` eager attack steps, then launches
geoa3_cell_sort + geoa3_knn_cells (hint = previous step's lists) once with the profiler on.  Never quote timings from
this script.   python tools/profile_knn_cells.py [--batch 250] [--steps 80] [--grid 8]"""
import argparse
import os.path as osp
import sys

import torch

ROOT = osp.dirname(osp.dirname(osp.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from geoa3_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=250)
ap.add_argument("--steps", type=int, default=80)
ap.add_argument("--kref", type=float, default=17)
ap.add_argument("--k", type=int, default=16)
a = ap.parse_args()
dev = torch.device("cuda", 0)
st, pins = bench.build_state("PointNet", a.batch, bench.NPTS, 0, a.batch, dev)
for _ in range(a.steps):
    st.step()
k, n = a.k, bench.NPTS
adv = (st.base + st.offset).detach().contiguous()
hint = st.hints.nbr[k].clone()
out = torch.empty_like(hint)
blobs = ops.cell_sort(adv, kref=a.kref)
ops.knn_cells(blobs, k + 1, drop=1, hint=hint, out=out)
ori = st.pc_ori.detach().contiguous()
bo = ops.cell_sort(ori, kref=4)
hb = st.hints
hj, hi = hb.jstar.clone(), hb.istar.clone()
outs = (torch.empty_like(hb.d1), torch.empty_like(hj), torch.empty_like(hb.d2), torch.empty_like(hi))
ops.nn_pair_cells(blobs, bo, hint_a2o=hj, hint_o2a=hi, out=outs)
torch.cuda.synchronize()
torch.cuda.profiler.start()
ops.cell_sort(adv, kref=a.kref, out=blobs)
ops.nn_pair_cells(blobs, bo, hint_a2o=hj, hint_o2a=hi, out=outs)
ops.knn_cells(blobs, k + 1, drop=1, hint=hint, out=out)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
