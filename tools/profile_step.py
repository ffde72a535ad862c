"""Three eager attack steps of the bench configuration (ncu target: launch list / per-kernel captures of ONE step;
never quote timings from this script).   python tools/profile_step.py [--batch B] [--arch PointNet]"""
import argparse
import os.path as osp
import sys

import torch

ROOT = osp.dirname(osp.dirname(osp.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=250)
ap.add_argument("--arch", default="PointNet")
ap.add_argument("--steps", type=int, default=3)
a = ap.parse_args()
dev = torch.device("cuda", 0)
st, pins = bench.build_state(a.arch, a.batch, bench.NPTS, 0, a.batch, dev)
for _ in range(a.steps - 1):
    st.step()
torch.cuda.synchronize()
torch.cuda.profiler.start()   # `ncu --profile-from-start off` captures the last step only
st.step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
