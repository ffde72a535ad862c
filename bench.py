#!/usr/bin/env python
"""bench.py — GeoA3 attack-iteration throughput on B200 (BASELINE.json metric, config[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--arch PointNet]

A "step" is ONE attack iteration (Attacker/geoA3_attack.py:238-368 of the reference) over a batch of
B=250 synthetic ModelNet-shaped instances (N=1024 points, random-init PointNet(40) in eval mode,
untargeted CE + 10*(1.0*CD + 0.1*HD + 1.0*curvature k=16), Adam lr 0.01): victim forward, fused
Chamfer/Hausdorff/curvature losses, backward to the offset, Adam update, success bookkeeping — replayed
as one CUDA graph.  Every GPU owns its own 250-instance batch (weak scaling, no data-path collective);
value = N_gpus * K / max-over-ranks device time.

Printed JSON (one line, rank 0): see the task contract; extra keys `roofline` (dominant own kernel,
measured live), `cpu_baseline` (oracle port of the same step on the host cores, bounded sample),
`loss_fwd_bwd_us` (per-kernel CUDA-event times of the loss path at the bench batch).
`--impl reference` times the CPU port of the reference path (kind "port") on all host cores.
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
CPU_SAMPLE_B = 4


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
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
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


def kernel_breakdown(pc_ori, nrm, adv, k):
    """CUDA-event time of each own kernel of the loss path at the bench batch (µs, mean of 10), in the exact
    configuration the attack step launches them: 1-NN seeded and kNN threshold hinted with the previous
    step's indices (here: the indices of a cloud one Adam step away)."""
    from geoa3_b200 import ops

    b, _, n = adv.shape
    prev = (adv - 0.003 * torch.sign(torch.randn_like(adv))).contiguous()
    _, hj, _, hi = ops.nn_pair(prev, pc_ori)
    hn = ops.knn(prev, prev, k + 1, drop=1)[0]
    perm, iperm = ops.visit_order(pc_ori)  # once per attack in the real driver (HintBuffers.ensure_order)
    nnkw = dict(hint_a2o=hj, hint_o2a=hi, perm_a=perm, perm_o=perm, iperm_a=iperm, iperm_o=iperm,
                ori_arranged=ops.arrange(pc_ori, perm))
    d1, js, d2, is_ = ops.nn_pair(adv, pc_ori, **nnkw)
    nbr = ops.knn(adv, adv, k + 1, drop=1, hint=hn)[0]
    nbr_o = ops.knn(pc_ori, pc_ori, k + 1, drop=1)[0]
    ko = ops.kappa_loss_fwd(pc_ori, normal=nrm, nbr=nbr_o)["kappa"]

    def fwd():
        return ops.kappa_loss_fwd(adv, normal=nrm, jstar=js, nbr=nbr, d_a2o=d1, d_o2a=d2, kappa_ori=ko, want_nrm=True,
                                  want_cd=True, want_hd=True, want_curv=True)

    out = fwd()
    g = torch.full((b,), 1.0 / b, device=adv.device)
    t = {
        "nn_pair": time_events(lambda: ops.nn_pair(adv, pc_ori, **nnkw), 10, 3),  # incl. the launch arranging adv
        "knn": time_events(lambda: ops.knn(adv, adv, k + 1, drop=1, hint=hn), 10, 3),
        "kappa_loss_fwd": time_events(fwd, 10, 3),
        "loss_bwd": time_events(lambda: ops.loss_bwd(adv, ori=pc_ori, nrm_adv=out["nrm"], kappa_adv=out["kappa"],
                                                     kappa_ori=ko, jstar=js, istar=is_, nbr=nbr, hd_arg=out["hd_arg"],
                                                     g_cd=g, g_hd=g, g_cu=g), 10, 3),
    }
    return {k_: round(v, 2) for k_, v in t.items()}


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


def cpu_baseline(steps, warmup, arch):
    """Oracle port of the same attack step on the host cores, bounded sample of CPU_SAMPLE_B instances."""
    from geoa3_b200.victims import PointNet
    from oracle import torch_port

    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    net = PointNet(40).eval()
    pc, nr, lab = make_inputs(CPU_SAMPLE_B, NPTS, 0)
    med, _ = torch_port.time_cpu_attack(net, torch.from_numpy(pc), torch.from_numpy(nr), torch.from_numpy(lab),
                                        steps=steps, warmup=warmup, k=KNN)
    per_instance_iter = med / CPU_SAMPLE_B
    value = 1.0 / (per_instance_iter * B_PER_GPU)
    return {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": "%d instances x %d timed iterations (median), PointNet + CD/HD/curvature + backward + Adam, dense "
                      "torch kNN as documented in the reference's comments; scaled to B=%d by instance count"
                      % (CPU_SAMPLE_B, steps, B_PER_GPU),
            "s_per_instance_iter": per_instance_iter}, med


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.time()
    cb, med = cpu_baseline(args.steps, max(args.warmup, 1), args.arch)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": med * 1e3 * B_PER_GPU / CPU_SAMPLE_B,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, 1), "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": round(time.time() - t0, 2)}
    print(json.dumps(line))


def workload_config(args, world):
    return {"workload": "GeoA3 attack iteration, %s N=%d, B=%d synthetic instances per GPU / 10 classes, CE(untargeted) + "
                        "10*(1.0*CD + 0.1*HD + 1.0*curvature k=%d), Adam lr 0.01 (BASELINE config[1])"
                        % (args.arch, NPTS, args.batch, KNN),
            "global_batch": args.batch * world, "npoint": NPTS, "parallelism": "instance-sharded x%d, no data-path collective" % world,
            "l2": "per-step working set (victim activations ~1 GB at B=250) exceeds the 126 MB L2; no explicit flush",
            "cuda_graph": not args.no_graph}


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    from geoa3_b200 import attack as atk
    from geoa3_b200 import dist as gdist
    from geoa3_b200 import ops
    from geoa3_b200.victims import build_victim

    rank, local_rank, world = gdist.init()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the GeoA3 hot path has no CPU fallback")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    b, n = args.batch, NPTS

    torch.manual_seed(0)
    net = build_victim(args.arch).to(dev).eval()
    for p in net.parameters():
        p.requires_grad_(False)
    pc_h, nr_h, lab_h = make_inputs(b, n, rank * b)
    pc_pin = torch.from_numpy(pc_h).pin_memory()
    nr_pin = torch.from_numpy(nr_h).pin_memory()
    off_pin = atk.default_offsets(b * world, n, 0, 0, gdist.shard_rows(b * world, world, rank)).pin_memory()
    pc_ori, nrm = pc_pin.to(dev), nr_pin.to(dev)
    target = torch.from_numpy(lab_h).to(dev)
    cfg = atk.make_cfg(attack_label="Untarget", curv_loss_knn=KNN)
    st = atk.AttackState(net, pc_ori, nrm, target, target, cfg, targeted=False, global_batch=b * world)
    st.begin_search_step(0, off_pin.to(dev))

    # count own kernel launches of one steady-state step (eager; the first step also pays the one-time arrangement
    # of the original cloud), then capture
    st.step()
    ops.LAUNCHES = 0
    st.step()
    launches_per_step = ops.LAUNCHES
    if not args.no_graph:
        st.capture()
    st.reset_global()
    st.begin_search_step(0, off_pin.to(dev))

    # ---------------- device-resident timing: W warm-up + K timed steps
    for _ in range(args.warmup):
        st.run_step()
    sampler = ClockSampler(local_rank)
    torch.cuda.synchronize()
    gdist.barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        st.run_step()
    e1.record()
    torch.cuda.synchronize()
    gdist.barrier()
    ms_total = gdist.max_over_ranks(e0.elapsed_time(e1), dev)
    clocks = sampler.stop()
    ms_per_step = ms_total / args.steps
    value = world * args.steps / (ms_total * 1e-3)
    final_loss = float(st.last["loss"].item())

    # ---------------- end-to-end: host buffers in, loss out, every step
    loss_host = torch.empty(b, dtype=torch.float32).pin_memory()
    from geoa3_b200 import loss_utils

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
    t0 = time.perf_counter()
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

    if rank == 0:
        # ---------------- roofline of the dominant own kernel, measured live (CUDA events, current stream)
        adv = (pc_ori + st.offset).detach().contiguous()
        kb = kernel_breakdown(pc_ori, nrm, adv, KNN)
        pk = peaks()
        top = max(kb, key=kb.get)
        alg_bytes = {"nn_pair": 40 * b * n, "knn": (12 + 4 * KNN) * b * n, "kappa_loss_fwd": (36 + 4 * KNN) * b * n,
                     "loss_bwd": (56 + 4 * KNN) * b * n}[top]
        alg_flop = {"nn_pair": 16.0 * b * n * n, "knn": 8.0 * b * n * n}.get(top, 0.0)
        # DRAM bytes per launch of each kernel from the committed `ncu --set full` captures (profiles/ncu_r1_summary.md:
        # dram__bytes_read.sum + dram__bytes_write.sum at this exact shape); outputs mostly stay in the 126 MB L2
        ncu_traffic = {"knn": 19.48e6, "nn_pair": 8.20e6, "kappa_loss_fwd": 26.66e6, "loss_bwd": 29.76e6}
        t_s = kb[top] * 1e-6
        achieved = alg_bytes / t_s / 1e9
        line["roofline"] = {"kernel": top, "bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
                            "frac": achieved / pk["hbm_gbs"], "traffic": ncu_traffic.get(top), "peak_source": pk["source"],
                            "algorithmic_bytes_per_launch": alg_bytes,
                            "note": "distance kernels are FP32-issue bound (SURVEY §8d); see fp32",
                            "fp32": {"achieved_tflops": alg_flop / t_s / 1e12, "peak_tflops": pk["fp32_tflops"],
                                     "frac": alg_flop / t_s / 1e12 / pk["fp32_tflops"],
                                     "peak_source": "ubench/fp32_peak.cu measured on this pool"}}
        line["loss_fwd_bwd_us"] = dict(kb, total=round(sum(kb.values()), 2),
                                       hbm_frac_of_52BN=round(52 * b * n / (sum(kb.values()) * 1e-6) / 1e9 / pk["hbm_gbs"], 5))
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"], _ = cpu_baseline(5, 1, args.arch)
        print(json.dumps(line))
    gdist.barrier()
    if torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
