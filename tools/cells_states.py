"""Saves / replays REAL attack states for A/B timing of the search kernels.
    python tools/cells_states.py --save /tmp/states.pt            (runs a 250-instance PointNet attack with the product lib)
    GEOA3_SO_PATH=variants/lib_x.so python tools/cells_states.py --load /tmp/states.pt [--kref 17] [--kori 4]
Replay times cell_sort / nn_pair_cells / knn_cells (CUDA events, L2 flushed) on every saved state and checks the
results against the saved product results (members as sets)."""
import argparse
import json
import os.path as osp
import sys

import torch

ROOT = osp.dirname(osp.dirname(osp.abspath(__file__)))
sys.path.insert(0, ROOT)

ap = argparse.ArgumentParser()
ap.add_argument("--save", type=str, default=None)
ap.add_argument("--load", type=str, default=None)
ap.add_argument("--batch", type=int, default=250)
ap.add_argument("--kref", type=float, default=17)
ap.add_argument("--kori", type=float, default=4)
ap.add_argument("--steps", type=str, default="20,150,499")
a = ap.parse_args()
k = 16
if a.save:
    import bench
    from geoa3_b200 import ops
    st, pins = bench.build_state("PointNet", a.batch, bench.NPTS, 0, a.batch, torch.device("cuda", 0))
    probe = {int(s) for s in a.steps.split(",")}
    out = []
    for step in range(max(probe) + 1):
        st.step()
        if step in probe:
            adv = (st.base + st.offset).detach().contiguous()
            hb = st.hints
            new = ops.knn(adv, adv, k + 1, drop=1)[0].sort(-1)[0]
            nn = ops.nn_pair(adv, st.pc_ori.detach().contiguous())
            out.append(dict(step=step, adv=adv.cpu(), ori=st.pc_ori.detach().cpu(), hint=hb.nbr[k].cpu(), hj=hb.jstar.cpu(),
                            hi=hb.istar.cpu(), want_nbr=new.cpu(), want_nn=[x.cpu() for x in nn]))
    torch.save(out, a.save)
    print("saved", len(out), "states")
else:
    from geoa3_b200 import ops
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def t(fn, iters=9):
        ts = []
        for _ in range(iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        return round(sorted(ts)[len(ts) // 2], 1)

    for s in torch.load(a.load):
        adv, ori, hint, hj, hi = (s[x].cuda() for x in ("adv", "ori", "hint", "hj", "hi"))
        ca, co = ops.cell_sort(adv, kref=a.kref), ops.cell_sort(ori, kref=a.kori)
        out = torch.empty_like(hint)
        r = dict(step=s["step"], sort=t(lambda: ops.cell_sort(adv, kref=a.kref, out=ca)),
                 nn=t(lambda: ops.nn_pair_cells(ca, co, hint_a2o=hj, hint_o2a=hi)),
                 knn=t(lambda: ops.knn_cells(ca, k + 1, drop=1, hint=hint, out=out)))
        r["knn_ok"] = bool(torch.equal(out.sort(-1)[0].cpu(), s["want_nbr"]))
        nn = ops.nn_pair_cells(ca, co, hint_a2o=hj, hint_o2a=hi)
        r["nn_ok"] = all(torch.equal(x.cpu(), y) for x, y in zip(nn, s["want_nn"]))
        print(json.dumps(r), flush=True)
