#!/usr/bin/env python
"""bench.py — GeoA3 attack-iteration throughput on B200 (BASELINE.json metric; headline = config[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--arch PointNet]

A "step" is ONE attack iteration (Attacker/geoA3_attack.py:238-368 of the reference) over a batch of
B=250 synthetic ModelNet-shaped instances (N=1024 points, random-init PointNet(40) in eval mode,
untargeted CE + 10*(1.0*CD + 0.1*HD + 1.0*curvature k=16), Adam lr 0.01): victim forward, fused
Chamfer/Hausdorff/curvature losses, backward to the offset, Adam update, success bookkeeping — replayed
as one CUDA graph.

Headline (`value`, unchanged since round 1): every GPU owns its own 250-instance batch (weak scaling, no
data-path collective); value = N_gpus * K / max-over-ranks device time.

Sub-records of the same JSON line (rank 0 prints ONE line):
  strong      config[1] as the north star states it: ONE 250-instance batch sharded by instance over the N
              ranks (geoa3_b200.dist), K steps per rank + the final NCCL all-gather of the per-instance
              statistics, device time max over ranks.
  strong_msg  config[4]: PointNet++ MSG, 2 000 instances sharded over the N ranks (micro-batches of 250 per
              rank, one captured step reused through AttackState.load_batch), + the final all-gather.
  configs     (N=1 only) config[2] PointNet++ SSG step with every own op timed beside the reference's kernel
              recompiled for sm_100 (oracle/_ref, baseline leg) and its HBM-roofline fraction; config[3] the
              curvature-loss sweep (B=64, N=1024/4096/10000, k=16/32).
  loss_fwd_bwd_us / gpu_reference_loss_us   own loss-path kernels vs the dense-torch loss the reference documents
              (Lib/loss_utils.py:30-31,54-56,67-69) on the same B200.
  roofline    dominant own kernel of the loss path, measured live.
  cpu_baseline  the oracle port of the same step on the host cores: ONE real step of all 250 instances
              (chunks of 25), plus config[0] (b=1, 10 iterations, loss-only and with PointNet).
`--impl reference` runs the CPU port for real at B=250 (reduced step count, reported as run).
"""
import argparse
import json
import os
import os.path as osp
import statistics
import subprocess
import sys
import threading
import time

ROOT = osp.dirname(osp.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "geoa3_attack_iters_per_s"
UNIT = "attack-iters/s (one iter = one Adam step of a 250-instance batch; whole job, all GPUs)"
B_PER_GPU, NPTS, KNN = 250, 1024, 16
MSG_TOTAL, MSG_MICRO = 2000, 250
CPU_CHUNK = 25


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--arch", default="PointNet", choices=["PointNet", "PointNetPP_ssg", "PointNetPP_msg"])
    ap.add_argument("--batch", type=int, default=B_PER_GPU)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="headline + strong only (skip configs[2..4], baselines)")
    ap.add_argument("--msg-total", type=int, default=MSG_TOTAL)
    ap.add_argument("--no-fold-bn", action="store_true", help="attack the victim as built (eval-mode BatchNorm kept as layers)")
    return ap.parse_args()


# ------------------------------------------------------------------ clocks sampler
class ClockSampler(object):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0=None, t1=None):
        """Summary of the samples taken inside the wall-clock window [t0, t1] (the timed region); nvidia-smi needs a
        few hundred ms to deliver its first sample, so the sampler is started well before the window opens."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = [r for (ts, r) in self.rows if t0 is None or (t0 - 0.05 <= ts <= t1 + 0.15)]
        for r in inside:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for nme, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------ workload
def make_inputs(b, n, start):
    from geoa3_b200 import synth

    base = min(b, 50)
    pc, nr, lab = synth.make_batch(base, n, start)
    reps = (b + base - 1) // base
    pc = np.tile(pc, (reps, 1, 1))[:b]
    nr = np.tile(nr, (reps, 1, 1))[:b]
    lab = np.tile(lab, reps)[:b]
    return pc, nr, lab


_FLUSH = None


def time_events(fn, iters, warm):
    """Mean CUDA-event time of fn() in µs.  A 256 MB memset is queued in front of every timed call: it flushes
    the 126 MB L2 and keeps the GPU busy while Python prepares the launch, so the interval between the two
    events is the kernel itself, not the host-side launch latency."""
    global _FLUSH
    if _FLUSH is None:
        _FLUSH = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        _FLUSH.zero_()
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return sum(ts) / len(ts)


def peaks():
    p = {"hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}
    f = osp.join(ROOT, "MEASURED_PEAKS.json")
    if osp.exists(f):
        try:
            p = {"hbm_gbs": float(json.load(open(f))["hbm_gbs"]), "source": "measured (MEASURED_PEAKS.json)"}
        except Exception:
            pass
    p["fp32_tflops"] = 71.6  # own measurement, ubench/fp32_peak.cu on this pool's B200 (profiles/fp32_peak_r1.jsonl)
    return p


def ncu_traffic():
    """DRAM bytes per launch of the own loss kernels from the committed `ncu --set full` capture of this round
    (profiles/ncu_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum at B=250, N=1024, k=16)."""
    f = osp.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        return json.load(open(f))
    except Exception:
        return {}


FOLD_BN = True


def build_state(arch, b, n, row0, global_batch, dev, rows=None, graph=True):
    """Victim + AttackState for rows [row0, row0+b) of a `global_batch`-row attack — exactly what the public
    attack() sets up: parameters frozen, eval-mode BatchNorm folded into the conv / linear weights."""
    from geoa3_b200 import attack as atk
    from geoa3_b200.victims import build_victim, fold_batchnorm

    torch.manual_seed(0)
    net = build_victim(arch).to(dev).eval()
    pc_h, nr_h, lab_h = make_inputs(b, n, row0)
    if FOLD_BN:
        net = fold_batchnorm(net, torch.from_numpy(pc_h[:2]).to(dev))
    for p in net.parameters():
        p.requires_grad_(False)
    pc_pin, nr_pin = torch.from_numpy(pc_h).pin_memory(), torch.from_numpy(nr_h).pin_memory()
    off_pin = atk.default_offsets(global_batch, n, 0, 0, rows if rows is not None else range(row0, row0 + b)).pin_memory()
    target = torch.from_numpy(lab_h).to(dev)
    cfg = atk.make_cfg(attack_label="Untarget", curv_loss_knn=KNN)
    st = atk.AttackState(net, pc_pin.to(dev), nr_pin.to(dev), target, target, cfg, targeted=False,
                         global_batch=global_batch)
    st.begin_search_step(0, off_pin.to(dev))
    return st, (pc_pin, nr_pin, off_pin)


def count_and_capture(st, off_dev, graph=True):
    from geoa3_b200 import ops

    st.step()          # the first step also pays the one-time arrangement of the original cloud
    ops.LAUNCHES = 0
    st.step()
    launches = ops.LAUNCHES
    if graph:
        st.capture()
    st.reset_global()
    st.begin_search_step(0, off_dev)
    return launches


def timed_steps(st, steps, warmup, dev, gdist, tail=None):
    """W warm-up + K timed steps (+ optional tail() inside the timed region); device time, max over ranks (ms)."""
    for _ in range(warmup):
        st.run_step()
    st.step_idx.zero_()  # (the per-step loss log has iter_max_steps rows: keep its row index inside it)
    torch.cuda.synchronize()
    gdist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(steps):
        st.run_step()
        if (i + 1) % 400 == 0:
            st.step_idx.zero_()
    out = tail() if tail is not None else None
    e1.record()
    torch.cuda.synchronize()
    gdist.barrier()
    return gdist.max_over_ranks(e0.elapsed_time(e1), dev), out


# ------------------------------------------------------------------ loss-path kernels
def kernel_breakdown(pc_ori, nrm, prev_adv, adv, k):
    """CUDA-event time of each own kernel of the loss path at the bench batch (µs, mean of 10), in the exact
    configuration the attack step launches them (loss_utils._GeoLoss with HintBuffers) and on the state a REAL attack
    produces: `prev_adv` and `adv` are the clouds of two consecutive Adam steps of the timed run, the search hints are
    the indices found on `prev_adv` (kept frozen, so every repetition sees the same previous-step hints)."""
    from geoa3_b200 import loss_utils as L

    b, _, n = adv.shape
    hb = L.HintBuffers()
    ko = L._get_kappa_ori(pc_ori, nrm, k).detach()
    L.step_plan(prev_adv, pc_ori, nrm, ko, k, hb)         # fills the hint buffers from the previous step's cloud
    hb.frozen = {"jstar": hb.jstar.clone(), "istar": hb.istar.clone(), "nbr": {k: hb.nbr[k].clone()}}
    plan = L.step_plan(adv, pc_ori, nrm, ko, k, hb)       # the launches of one steady-state step, as closures
    g = torch.full((b,), 1.0 / b, device=adv.device)
    t = {}
    for name, fn in plan["launches"](g):
        t[name] = time_events(fn, 10, 3)
    return {k_: round(v, 2) for k_, v in t.items()}


def dense_torch_loss_us(pc_ori, nrm, adv, k):
    """The GPU "reference kernel" bar of BASELINE.md section 3(i): the dense-torch formulation the reference documents
    next to every pytorch3d call (Lib/loss_utils.py:30-31,54-56,67-69: [b,n,n] squared distances + topk), forward +
    backward of CD + 0.1*HD + curvature on the same B200, plain torch ops (restated here; nothing from oracle/)."""
    b, _, n = adv.shape
    chunk = 50  # [chunk,n,n] matrices: the full batch at once would need > 10 GB of temporaries

    def knn_dense(p1, p2, K):
        d = ((p1.unsqueeze(3) - p2.unsqueeze(2)) ** 2).sum(1)
        return torch.topk(d, K, dim=2, largest=False, sorted=True)

    def kappa(pc, normal):
        bb = pc.shape[0]
        idx = knn_dense(pc, pc, k + 1)[1][:, :, 1:].contiguous()
        nn_pts = torch.gather(pc, 2, idx.view(bb, 1, n * k).expand(bb, 3, n * k)).view(bb, 3, n, k)
        v = nn_pts - pc.unsqueeze(3)
        v = v / v.norm(2, 1, keepdim=True).clamp(min=1e-12)
        return torch.abs((v * normal.unsqueeze(3)).sum(1)).mean(2)

    ko = torch.cat([kappa(pc_ori[i:i + chunk], nrm[i:i + chunk]) for i in range(0, b, chunk)])

    def run():
        for i in range(0, b, chunk):
            a = adv[i:i + chunk].detach().requires_grad_(True)
            o, nn_, kk = pc_ori[i:i + chunk], nrm[i:i + chunk], ko[i:i + chunk]
            bb = a.shape[0]
            d1, j1 = knn_dense(a, o, 1)
            d2, _ = knn_dense(o, a, 1)
            cd = d1.squeeze(-1).mean(-1) + d2.squeeze(-1).mean(-1)
            hd = d1.squeeze(-1).max(-1)[0]
            normal = torch.gather(nn_, 2, j1.view(bb, 1, n).expand(bb, 3, n))
            cu = ((kappa(a, normal) - torch.gather(kk, 1, j1.squeeze(-1))) ** 2).mean(-1)
            (cd + 0.1 * hd + cu).sum().backward()

    return round(time_events(run, 3, 1), 1)


# ------------------------------------------------------------------ config[2]: PointNet++ SSG ops
def ssg_ops_record(b, dev):
    """Own pointnet2 ops at the SSG shapes (Model/PointNetPP_ssg.py:65-87) with their HBM-roofline fraction, and the
    reference's kernels recompiled for sm_100 (oracle/_ref — baseline leg only) timed beside them."""
    from geoa3_b200 import ops

    pk = peaks()
    pc_h, _, _ = make_inputs(b, NPTS, 0)
    ori = torch.from_numpy(pc_h).to(dev)
    xyz = ori.transpose(1, 2).contiguous()
    fi = ops.furthest_point_sampling(xyz, 512)
    new = ops.gather_points(ori, fi).transpose(1, 2).contiguous()
    idx = ops.ball_query(new, xyz, 0.2, 64)
    fi2 = ops.furthest_point_sampling(new, 128)
    new2 = ops.gather_points(new.transpose(1, 2).contiguous(), fi2).transpose(1, 2).contiguous()
    idx2 = ops.ball_query(new2, new, 0.4, 64)
    feats = torch.randn(b, 128, 512, device=dev)
    go = torch.randn(b, 128, 128, 64, device=dev)
    go3 = torch.randn(b, 3, 512, 64, device=dev)
    n, m1, m2, ns = NPTS, 512, 128, 64

    def grp(c, nn_, m):  # 4*C*n + 4*m*ns + 4*C*m*ns per cloud
        return b * (4 * c * nn_ + 4 * m * ns + 4 * c * m * ns)

    cases = [
        ("fps_1024_512", lambda e: e.furthest_point_sampling(xyz, 512), b * (12 * n + 4 * m1)),
        ("fps_512_128", lambda e: e.furthest_point_sampling(new, 128), b * (12 * m1 + 4 * m2)),
        ("gather_3_512", lambda e: e.gather_points(ori, fi), b * 28 * m1),
        ("ball_query_r.2_ns64", lambda e: e.ball_query(new, xyz, 0.2, 64), b * (12 * m1 + 12 * n + 4 * m1 * ns)),
        ("ball_query_r.4_ns64", lambda e: e.ball_query(new2, new, 0.4, 64), b * (12 * m2 + 12 * m1 + 4 * m2 * ns)),
        ("group_c3_m512", lambda e: e.group_points(ori, idx), grp(3, n, m1)),
        ("group_c128_m128", lambda e: e.group_points(feats, idx2), grp(128, m1, m2)),
        ("group_grad_c128_m128", lambda e: e.group_points_grad(go, idx2, 512), grp(128, m1, m2)),
        ("group_grad_c3_m512", lambda e: e.group_points_grad(go3, idx, 1024), grp(3, n, m1)),
    ]
    ext = None
    try:
        from oracle import build_ref  # baseline leg: the reference's own kernels, never on the product path

        ext = build_ref.load_ref()
    except Exception:
        ext = None
    rec = {}
    for name, fn, byts in cases:
        us = time_events(lambda: fn(ops), 10, 3)
        r = {"us": round(us, 2), "algorithmic_bytes": byts, "hbm_frac": round(byts / (us * 1e-6) / 1e9 / pk["hbm_gbs"], 4)}
        if ext is not None:
            try:
                r["reference_kernel_sm100_us"] = round(time_events(lambda: fn(ext), 5, 2), 2)
                r["speedup_vs_reference_kernel"] = round(r["reference_kernel_sm100_us"] / us, 2)
            except Exception as e:  # noqa
                r["reference_kernel_error"] = str(e)[:80]
        rec[name] = r
    return rec


def sweep_record(dev):
    """config[3]: kNN + kappa kernels, B=64, N in {1024,4096,10000}, k in {16,32}, hinted like inside the attack."""
    from geoa3_b200 import ops, synth

    pk = peaks()
    B, rows = 64, []
    for n in (1024, 4096, 10000):
        pc, nr, _ = synth.make_batch(16, n)
        ori = torch.from_numpy(np.tile(pc, (4, 1, 1))).to(dev)
        nrm = torch.from_numpy(np.tile(nr, (4, 1, 1))).to(dev)
        adv = ori + torch.from_numpy(synth.make_offsets(B, n)).to(dev)
        prev = (adv - 0.003 * torch.sign(torch.randn_like(adv))).contiguous()
        _, js, _, _ = ops.nn_pair(adv, ori)
        pm, ipm = ops.visit_order(ori)
        for k in (16, 32):
            hn = ops.knn(prev, prev, k + 1, drop=1)[0]
            t_plain = time_events(lambda: ops.knn(adv, adv, k + 1, drop=1), 3, 1)
            t_hint = time_events(lambda: ops.knn(adv, adv, k + 1, drop=1, hint=hn), 3, 1)
            t_prune = time_events(lambda: ops.knn(adv, adv, k + 1, drop=1, hint=hn, perm_q=pm, perm_c=pm, iperm_c=ipm), 3, 1)
            t_set = time_events(lambda: ops.knn(adv, adv, k + 1, drop=1, hint=hn, perm_q=pm, perm_c=pm, iperm_c=ipm,
                                                members_only=True), 3, 1)
            t_cells = t_sort = None
            if n <= 4096:   # the cell-grid search (the attack step uses it up to HintBuffers.CELLS_MAX_N points)
                cells = ops.cell_sort(adv, kref=k + 1)
                t_sort = time_events(lambda: ops.cell_sort(adv, kref=k + 1, out=cells), 3, 1)
                t_cells = time_events(lambda: ops.knn_cells(cells, k + 1, drop=1, hint=hn), 3, 1)
            nbr = ops.knn(adv, adv, k + 1, drop=1, hint=hn)[0]
            t_kap = time_events(lambda: ops.kappa_loss_fwd(adv, normal=nrm, jstar=js, nbr=nbr), 3, 1)
            best = min(t_hint, t_prune, t_set, (t_cells + t_sort) if t_cells is not None else 1e30)
            byts = (28 + 4 * k) * B * n
            rows.append({"n": n, "k": k, "knn_us": round(t_plain, 1), "knn_hinted_us": round(t_hint, 1),
                         "knn_hinted_pruned_us": round(t_prune, 1), "knn_members_hinted_pruned_us": round(t_set, 1),
                         "knn_cells_us": None if t_cells is None else round(t_cells, 1),
                         "cell_sort_us": None if t_sort is None else round(t_sort, 1),
                         "kappa_us": round(t_kap, 1),
                         "hbm_frac": round(byts / ((best + t_kap) * 1e-6) / 1e9 / pk["hbm_gbs"], 5),
                         "fp32_tflops": round(8.0 * B * n * n / (best * 1e-6) / 1e12, 2)})
    return rows


# ------------------------------------------------------------------ CPU arm
def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


class CpuBatch(object):
    """The oracle port of the attack step on the host cores, the 250-instance batch walked in chunks of 25 (a dense
    [b,n,n] formulation of all 250 instances at once would need tens of GB; the reference's own CLI default is
    batch_size 2, main_attack.py:326).  One step() = one real attack iteration of ALL instances."""

    def __init__(self, b, with_net=True):
        from geoa3_b200.victims import PointNet
        from oracle import torch_port

        torch.set_num_threads(os.cpu_count() or 1)
        torch.manual_seed(0)
        net = PointNet(40).eval()
        for p in net.parameters():
            p.requires_grad_(False)
        pc, nr, lab = make_inputs(b, NPTS, 0)
        self.with_net = with_net
        self.chunks = [torch_port.CpuAttackStep(net, torch.from_numpy(pc[i:i + CPU_CHUNK]), torch.from_numpy(nr[i:i + CPU_CHUNK]),
                                                torch.from_numpy(lab[i:i + CPU_CHUNK]), KNN)
                       for i in range(0, b, CPU_CHUNK)]

    def step(self, only_first=False):
        for c in (self.chunks[:1] if only_first else self.chunks):
            c.step(self.with_net)


def cpu_baseline_record(b):
    """Bounded sample on the host cores: ONE real step of all `b` instances (after a one-chunk warm-up), plus
    BASELINE config[0]: b=1, 10 iterations, loss-only and with PointNet."""
    from geoa3_b200.victims import PointNet
    from oracle import torch_port

    job = CpuBatch(b)
    job.step(only_first=True)
    t0 = time.perf_counter()
    job.step()
    s_step = time.perf_counter() - t0
    torch.manual_seed(0)
    net = PointNet(40).eval()
    pc, nr, lab = make_inputs(1, NPTS, 0)
    args1 = (net, torch.from_numpy(pc), torch.from_numpy(nr), torch.from_numpy(lab))
    med_full, _ = torch_port.time_cpu_attack(*args1, steps=10, warmup=2, k=KNN, with_net=True)
    med_loss, _ = torch_port.time_cpu_attack(*args1, steps=10, warmup=2, k=KNN, with_net=False)
    return {"value": 1.0 / s_step, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "cpu_model": cpu_model(),
            "sample": "1 timed attack iteration of all %d instances (chunks of %d; PointNet + CD/HD/curvature + backward + "
                      "Adam, dense torch kNN as documented in the reference's comments) after a one-chunk warm-up"
                      % (b, CPU_CHUNK),
            "s_per_step": round(s_step, 3),
            "config0_b1_10iters": {"loss_only_fwd_bwd_adam_ms": round(med_loss * 1e3, 2),
                                   "with_pointnet_ms": round(med_full * 1e3, 2), "stat": "median of 10"}}


def run_reference(args):
    """--impl reference: the CPU port for real at the headline config (all 250 instances per step), reduced step
    count so the run ends within a few minutes; steps / warmup / ms_per_step are reported AS RUN."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.time()
    steps, warm = max(1, min(args.steps, 3)), max(0, min(args.warmup, 1))
    job = CpuBatch(args.batch)
    for _ in range(warm):
        job.step()
    ts = []
    for _ in range(steps):
        t1 = time.perf_counter()
        job.step()
        ts.append(time.perf_counter() - t1)
    total = sum(ts)
    value = steps / total
    cb = {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "cpu_model": cpu_model(),
          "sample": "%d timed + %d warm-up attack iterations of all %d instances (chunks of %d), nothing extrapolated"
                    % (steps, warm, args.batch, CPU_CHUNK)}
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "requested": {"steps": args.steps, "warmup": args.warmup},
            "extrapolated": False, "ms_per_step": total / steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, 1), "cpu_baseline": cb,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": round(time.time() - t0, 2)}
    print(json.dumps(line))


def workload_config(args, world):
    return {"workload": "GeoA3 attack iteration, %s N=%d, B=%d synthetic instances per GPU / 10 classes, CE(untargeted) + "
                        "10*(1.0*CD + 0.1*HD + 1.0*curvature k=%d), Adam lr 0.01 (BASELINE config[1])"
                        % (args.arch, NPTS, args.batch, KNN),
            "global_batch": args.batch * world, "npoint": NPTS, "parallelism": "instance-sharded x%d, no data-path collective" % world,
            "l2": "per-step working set (victim activations ~1 GB at B=250) exceeds the 126 MB L2; no explicit flush",
            "cuda_graph": not args.no_graph,
            "victim": "random-init, eval mode, parameters frozen, BatchNorm folded into conv/linear weights (exact algebra)"
                      if not args.no_fold_bn else "random-init, eval mode, parameters frozen"}


# ------------------------------------------------------------------ strong-scaling records
def strong_record(args, rank, world, dev, gdist):
    """config[1] as the north star states it: ONE `args.batch`-instance PointNet batch sharded over the ranks,
    K steps + pack_stats + the final all-gather inside the timed region."""
    total = args.batch
    rows = gdist.shard_rows(total, world, rank)
    st, pins = build_state("PointNet", len(rows), NPTS, rows.start, total, dev, rows=rows)
    off_dev = pins[2].to(dev)
    count_and_capture(st, off_dev, graph=not args.no_graph)

    def tail():
        return gdist.gather_stats(gdist.state_stats(st), total)

    tail()  # warm the collective (NCCL communicator set-up is not part of an attack's steady state)
    ms, stats = timed_steps(st, args.steps, args.warmup, dev, gdist, tail)
    # the all-gather alone
    torch.cuda.synchronize()
    gdist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    tail()
    e1.record()
    torch.cuda.synchronize()
    ag_ms = gdist.max_over_ranks(e0.elapsed_time(e1), dev)
    rec = {"workload": "ONE %d-instance PointNet attack batch sharded by instance over %d rank(s): %d steps per rank + "
                       "final all-gather of the [%d,%d] per-instance statistics (NCCL), device time max over ranks"
                       % (total, world, args.steps, total, len(gdist.STAT_FIELDS)),
           "scaling": "strong", "global_batch": total, "rows_per_rank": [len(gdist.shard_rows(total, world, r)) for r in range(world)],
           "steps": args.steps, "ms_total": ms, "ms_per_step": ms / args.steps, "value": args.steps / (ms * 1e-3),
           "unit": "attack-iters/s of the %d-instance batch" % total, "allgather_ms": ag_ms,
           "stats_rows_gathered": int(stats.shape[0])}
    del st
    torch.cuda.empty_cache()
    return rec


def msg_record(args, rank, world, dev, gdist):
    """config[4]: PointNet++ MSG attack on `--msg-total` instances sharded over the ranks; every rank walks its
    rows in micro-batches of <= 250 through ONE captured step (AttackState.load_batch), then the all-gather."""
    from geoa3_b200 import ops

    total = args.msg_total
    rows = gdist.shard_rows(total, world, rank)
    micro = min(MSG_MICRO, len(rows))
    nmb = (len(rows) + micro - 1) // micro
    steps = max(1, min(args.steps, 3))
    st, pins = build_state("PointNetPP_msg", micro, NPTS, rows.start, total, dev, rows=range(rows.start, rows.start + micro))
    off_dev = pins[2].to(dev)
    launches = count_and_capture(st, off_dev, graph=not args.no_graph)
    batches = []
    for i in range(nmb):  # device-resident inputs of every micro-batch (the last one is padded by wrapping around)
        r0 = rows.start + min(i * micro, len(rows) - micro)
        pc_h, nr_h, lab_h = make_inputs(micro, NPTS, r0)
        batches.append((torch.from_numpy(pc_h).to(dev), torch.from_numpy(nr_h).to(dev), torch.from_numpy(lab_h).to(dev)))
    blocks = []

    def job():
        del blocks[:]
        for pc_d, nr_d, lab_d in batches:
            st.load_batch(pc_d, nr_d, lab_d)
            st.begin_search_step(0, off_dev)
            for _ in range(steps):
                st.run_step()
            blocks.append(gdist.state_stats(st))
        local = torch.cat(blocks, 0)[: len(rows)]
        return gdist.gather_stats(local, total)

    job()  # warm-up pass (also warms the collective)
    torch.cuda.synchronize()
    gdist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    stats = job()
    e1.record()
    torch.cuda.synchronize()
    gdist.barrier()
    ms = gdist.max_over_ranks(e0.elapsed_time(e1), dev)
    rec = {"workload": "PointNet++ MSG N=%d, %d instances sharded over %d rank(s), micro-batches of %d per rank through one "
                       "captured step, %d attack iterations per instance + final all-gather (BASELINE config[4])"
                       % (NPTS, total, world, micro, steps),
           "scaling": "strong", "global_instances": total, "micro_batches_per_rank": nmb, "steps_per_instance": steps,
           "ms_total": ms, "ms_per_microbatch_step": ms / (nmb * steps),
           "value": total * steps / (ms * 1e-3), "unit": "instance-iterations/s (whole job, all GPUs)",
           "own_launches_per_step": launches, "stats_rows_gathered": int(stats.shape[0])}
    del st, batches
    torch.cuda.empty_cache()
    return rec


def ssg_step_record(args, dev, gdist):
    """config[2]: PointNet++ SSG attack step at B=250 (own FPS / ball_query / group_points every step)."""
    st, pins = build_state("PointNetPP_ssg", args.batch, NPTS, 0, args.batch, dev)
    launches = count_and_capture(st, pins[2].to(dev), graph=not args.no_graph)
    steps = max(3, min(args.steps, 10))
    ms, _ = timed_steps(st, steps, 2, dev, gdist)
    del st
    torch.cuda.empty_cache()
    return {"workload": "PointNet++ SSG N=%d attack iteration, B=%d (BASELINE config[2])" % (NPTS, args.batch),
            "ms_per_step": ms / steps, "value": steps / (ms * 1e-3), "unit": UNIT, "steps": steps,
            "own_launches_per_step": launches}


def headline(args, rank, local_rank, world, dev, gdist):
    """The unchanged headline: device-resident steps (value) and the host-buffer e2e loop of config[1]."""
    from geoa3_b200 import loss_utils

    b, n = args.batch, NPTS
    sampler = ClockSampler(local_rank)
    sampler.start()
    st, (pc_pin, nr_pin, off_pin) = build_state(args.arch, b, n, rank * b, b * world, dev)
    pc_ori, nrm = st.pc_ori, st.normal_ori
    launches_per_step = count_and_capture(st, off_pin.to(dev), graph=not args.no_graph)

    # ---------------- device-resident timing: W warm-up + K timed steps
    for _ in range(args.warmup):
        st.run_step()
    torch.cuda.synchronize()
    w0 = time.time()
    ms_total, _ = timed_steps(st, args.steps, 0, dev, gdist)
    w1 = time.time()
    # the clock record covers the timed region; short regions are extended by untimed steps so nvidia-smi (100 ms
    # period) gets at least a few samples under the same load
    st.step_idx.zero_()
    while time.time() - w0 < 0.6:
        st.run_step()
        torch.cuda.synchronize()
        w1 = time.time()
    st.step_idx.zero_()
    clocks = sampler.stop(w0, w1)
    # two consecutive clouds of this run (>= 60 Adam steps in: past the first steps, where every point still moves by
    # the full learning rate) for the per-kernel timings below
    for _ in range(max(0, 60 - args.warmup - args.steps)):
        st.run_step()
    prev_adv = (pc_ori + st.offset).detach().clone()
    st.run_step()
    adv = (pc_ori + st.offset).detach().clone()
    steps_in = int(st.step_idx.item())
    st.step_idx.zero_()
    ms_per_step = ms_total / args.steps
    value = world * args.steps / (ms_total * 1e-3)
    final_loss = float(st.last["loss"].item())

    # ---------------- end-to-end: host buffers in, loss out, every step
    loss_host = torch.empty(b, dtype=torch.float32).pin_memory()
    # Double-buffered input path: every step's inputs come from pinned host memory (H2D inside the timed region,
    # one set per step), but the copy for step i+1 runs on a copy stream WHILE step i computes; the step itself
    # only does a device-to-device move out of the staging buffers.
    copy_stream = torch.cuda.Stream()
    stage = [torch.empty_like(pc_ori), torch.empty_like(nrm), torch.empty_like(st.offset)]
    staged, consumed = torch.cuda.Event(), torch.cuda.Event()

    def prefetch():
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed)  # the previous contents have been moved out
            stage[0].copy_(pc_pin, non_blocking=True)
            stage[1].copy_(nr_pin, non_blocking=True)
            stage[2].copy_(off_pin, non_blocking=True)
            staged.record(copy_stream)

    def e2e_step():
        cur = torch.cuda.current_stream()
        cur.wait_event(staged)
        with torch.no_grad():
            pc_ori.copy_(stage[0]); nrm.copy_(stage[1]); st.offset.copy_(stage[2])
            consumed.record(cur)
            prefetch()  # next step's H2D, overlapped with this step's compute
            st.kappa_ori.copy_(loss_utils._get_kappa_ori(pc_ori, nrm, KNN))  # inputs are new => recompute
        st.run_step()
        loss_host.copy_(st.loss_log[0], non_blocking=True)
        cur.synchronize()

    consumed.record(torch.cuda.current_stream())
    prefetch()
    for _ in range(min(3, args.warmup)):
        e2e_step()
    st.step_idx.zero_()
    torch.cuda.synchronize()
    gdist.barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    e2e_steps = max(5, min(args.steps, 50))
    for _ in range(e2e_steps):
        st.step_idx.zero_()
        e2e_step()
    f1.record()
    torch.cuda.synchronize()
    gdist.barrier()
    e2e_ms = gdist.max_over_ranks(f0.elapsed_time(f1), dev)
    e2e_value = world * e2e_steps / (e2e_ms * 1e-3)
    h2d = pc_pin.numel() * 4 + nr_pin.numel() * 4 + off_pin.numel() * 4
    d2h = loss_host.numel() * 4

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(args, world), "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps},
            "gpu_launches": launches_per_step * args.steps,
            "instance_iters_per_s": value * b, "final_loss": final_loss}

    line["loss_kernels_timed_at_step"] = steps_in
    return line, (prev_adv, adv), pc_ori.clone(), nrm.clone()



def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    global FOLD_BN
    FOLD_BN = not args.no_fold_bn

    from geoa3_b200 import dist as gdist
    from geoa3_b200 import loss_utils

    rank, local_rank, world = gdist.init()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the GeoA3 hot path has no CPU fallback")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    b, n = args.batch, NPTS

    line, adv, pc_keep, nrm_keep = headline(args, rank, local_rank, world, dev, gdist)
    loss_utils.clear_cache()
    import gc
    gc.collect()
    torch.cuda.empty_cache()

    def guarded(name, fn):
        try:
            line[name] = fn()
        except Exception as e:  # a failing sub-record must not cost the headline
            line[name] = {"error": "%s: %s" % (type(e).__name__, str(e)[:200])}

    # ---------------- strong scaling (every N): config[1] sharded + final all-gather; config[4] MSG 2000 instances
    guarded("strong", lambda: strong_record(args, rank, world, dev, gdist))
    if not args.no_extras:
        guarded("strong_msg", lambda: msg_record(args, rank, world, dev, gdist))

    if rank == 0:
        # ---------------- roofline of the dominant own kernel, measured live (CUDA events, current stream)
        kb = kernel_breakdown(pc_keep, nrm_keep, adv[0], adv[1], KNN)
        pk = peaks()
        top = max(kb, key=kb.get)
        alg_bytes = {"arrange": 28 * b * n, "cell_sort": 34 * b * n, "nn_pair": 40 * b * n, "knn": (12 + 4 * KNN) * b * n,
                     "knn_kappa": (12 + 4 * KNN + 24) * b * n, "kappa_loss_fwd": (36 + 4 * KNN) * b * n,
                     "loss_reduce": 20 * b * n, "loss_bwd": (56 + 4 * KNN) * b * n,
                     "geo_fwd_bwd": (84 + 4 * KNN) * b * n}.get(top, 52 * b * n)
        alg_flop = {"nn_pair": 16.0 * b * n * n, "knn": 8.0 * b * n * n, "knn_kappa": 8.0 * b * n * n}.get(top, 0.0)
        t_s = kb[top] * 1e-6
        achieved = alg_bytes / t_s / 1e9
        line["roofline"] = {"kernel": top, "bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
                            "frac": achieved / pk["hbm_gbs"], "traffic": ncu_traffic().get(top), "peak_source": pk["source"],
                            "algorithmic_bytes_per_launch": alg_bytes,
                            "note": "the search kernels are instruction-issue / shared-memory-latency bound, not HBM bound "
                                    "(SURVEY §8d); fp32 = flops of the full n x n scan the kernel REPLACES / its time (the "
                                    "cell walk evaluates ~5 % of those pairs)",
                            "fp32": {"achieved_tflops": alg_flop / t_s / 1e12, "peak_tflops": pk["fp32_tflops"],
                                     "frac": alg_flop / t_s / 1e12 / pk["fp32_tflops"],
                                     "peak_source": "ubench/fp32_peak.cu measured on this pool"}}
        line["loss_fwd_bwd_us"] = dict(kb, total=round(sum(kb.values()), 2),
                                       hbm_frac_of_52BN=round(52 * b * n / (sum(kb.values()) * 1e-6) / 1e9 / pk["hbm_gbs"], 5))
        if world == 1 and not args.no_extras:
            guarded("gpu_reference_loss_us", lambda: {
                "dense_torch_loss_fwd_bwd_us": dense_torch_loss_us(pc_keep, nrm_keep, adv[1], KNN),
                "what": "CD + 0.1*HD + curvature(k=16) forward+backward at B=%d, N=%d with the dense [b,n,n] + topk "
                        "formulation of Lib/loss_utils.py:30-31,54-56,67-69 in plain torch on this B200" % (b, n)})
            torch.cuda.empty_cache()
            cfgs = {}
            line["configs"] = cfgs
            for name, fn in (("2_pointnetpp_ssg_step", lambda: ssg_step_record(args, dev, gdist)),
                             ("2_pointnetpp_ssg_ops", lambda: ssg_ops_record(b, dev)),
                             ("3_curvature_sweep_B64", lambda: sweep_record(dev))):
                try:
                    cfgs[name] = fn()
                except Exception as e:
                    cfgs[name] = {"error": "%s: %s" % (type(e).__name__, str(e)[:200])}
                torch.cuda.empty_cache()
        if world == 1 and not args.no_cpu_baseline:
            guarded("cpu_baseline", lambda: cpu_baseline_record(b))
        print(json.dumps(line))
    gdist.barrier()
    if torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
