"""bench.py's config[3] record alone (curvature sweep: kNN + kappa kernels, B=64, N in {1024, 4096, 10000}, k in {16, 32}).
    python tools/sweep_only.py  -> one JSON list"""
import json
import os.path as osp
import sys

import torch

sys.path.insert(0, osp.dirname(osp.dirname(osp.abspath(__file__))))
import bench  # noqa: E402

print(json.dumps(bench.sweep_record(torch.device("cuda", 0))))
