"""ms per captured attack step at a given batch (PointNet, N=1024), with / without the side-stream overlap of the
geometry losses.   python tools/step_time.py [--batch 32]"""
import argparse
import json
import os.path as osp
import sys

import torch

sys.path.insert(0, osp.dirname(osp.dirname(osp.abspath(__file__))))
import bench  # noqa: E402
from geoa3_b200 import attack as atk  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
a = ap.parse_args()
dev = torch.device("cuda", 0)
for overlap in (True, False, True, False):
    atk.OVERLAP_GEO = overlap
    st, pins = bench.build_state("PointNet", a.batch, bench.NPTS, 0, 250, dev)
    st.capture()
    for _ in range(20):
        st.run_step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        st.run_step()
    e1.record()
    torch.cuda.synchronize()
    print(json.dumps(dict(batch=a.batch, overlap=overlap, graph=st.graph is not None, ms_per_step=round(e0.elapsed_time(e1) / 200, 4))), flush=True)
