"""Instance sharding of the attack across the GPUs of one box.

Each row of the attack batch is an independent optimisation (own offset, Adam state, scale_const,
bounds — Attacker/geoA3_attack.py:221-228,265-275; victims are in eval mode), so rank r simply owns a
contiguous block of instances: no collective inside the 10x500-step loop.  The only exchange is one
all-gather of a small per-instance statistics block at the very end (NCCL on GPUs, gloo in the CPU
tests).  One process per GPU, launched by torchrun; RANK / LOCAL_RANK / WORLD_SIZE come from the env.
"""
import os

import torch
import torch.distributed as dist

STAT_FIELDS = ("success", "best_loss", "best_step", "l2_offset", "linf_offset", "cd", "hd", "curv")


def env_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init(backend=None):
    """Initialises torch.distributed from the torchrun environment (no-op for world size 1)."""
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, local_rank, world


def shard_rows(total, world, rank):
    """Contiguous block of ceil(total/world) instances per rank (250 over 8 -> 32,...,32,26)."""
    per = (total + world - 1) // world
    lo = min(total, rank * per)
    return range(lo, min(total, lo + per))


def pack_stats(success, best_loss, best_step, best_attack, pc_ori, cd=None, hd=None, curv=None):
    """[local_b, F] float32 block of per-instance results (field order = STAT_FIELDS)."""
    off = (best_attack - pc_ori)
    z = torch.zeros_like(best_loss)
    cols = [success.to(torch.float32), best_loss.to(torch.float32), best_step.to(torch.float32),
            off.pow(2).sum((1, 2)).sqrt(), off.abs().amax((1, 2)),
            cd if cd is not None else z, hd if hd is not None else z, curv if curv is not None else z]
    return torch.stack(cols, 1).contiguous()


def gather_stats(local, total):
    """all_gather of the padded per-rank blocks -> [total, F] on every rank (the ONLY collective of the run)."""
    rank, _, world = env_world()
    if world == 1 or not dist.is_initialized():
        return local[:total]
    per = (total + world - 1) // world
    padded = torch.zeros(per, local.size(1), device=local.device, dtype=local.dtype)
    padded[: local.size(0)] = local
    out = torch.empty(world * per, local.size(1), device=local.device, dtype=local.dtype)
    dist.all_gather_into_tensor(out, padded)
    rows = []
    for r in range(world):
        sr = shard_rows(total, world, r)
        rows.append(out[r * per: r * per + len(sr)])
    return torch.cat(rows, 0)


def gather_clouds(local, total):
    """Optional second exchange: all_gather of the best adversarial clouds [local_b,3,n] -> [total,3,n]."""
    rank, _, world = env_world()
    if world == 1 or not dist.is_initialized():
        return local[:total]
    per = (total + world - 1) // world
    padded = torch.zeros((per,) + tuple(local.shape[1:]), device=local.device, dtype=local.dtype)
    padded[: local.size(0)] = local
    out = torch.empty((world * per,) + tuple(local.shape[1:]), device=local.device, dtype=local.dtype)
    dist.all_gather_into_tensor(out, padded)
    return torch.cat([out[r * per: r * per + len(shard_rows(total, world, r))] for r in range(world)], 0)


def shard_input(input_data, rows):
    """Rows `rows` of an attack batch.  `input_data` is the reference loader's item list
    [pcs [bs,l,n,3]|[bs,l,3,n], normals (same), gt_labels [bs,l], (target_labels [bs,l])]
    (Provider/modelnet10_instance250.py:82); the attack flattens (bs,l) into b = bs*l independent rows
    (Attacker/geoA3_attack.py:196-214), so a shard is rows of that flattened batch, returned as [len(rows),1,...]."""
    lo, hi = (rows.start, rows.stop) if isinstance(rows, range) else (rows[0], rows[-1] + 1)
    out = []
    for t in input_data:
        flat = t.reshape((t.shape[0] * t.shape[1],) + tuple(t.shape[2:]))
        out.append(flat[lo:hi].unsqueeze(1))
    return out


def state_stats(st):
    """[local_b, F] statistics block of a finished AttackState (device tensors only, no host sync)."""
    last = st.last
    return pack_stats(st.best_loss < 1e10, st.best_loss, st.best_attack_step, st.best_attack, st.pc_ori,
                      cd=last.get("dis"), hd=last.get("hd"), curv=last.get("curv"))


def attack_sharded(net, input_data, cfg, seed=0, ref_quirks=False, use_cuda_graph=True, with_clouds=False):
    """The whole multi-GPU attack of one batch: this rank takes its contiguous block of the
    b = bs*l rows, runs `attack()` on it with the GLOBAL batch size as loss divisor and the globally drawn initial
    offsets sliced to its rows (so every row follows the trajectory it has in the unsharded run), and the ranks
    exchange the per-instance statistics with ONE all_gather at the very end (NCCL over NVLink on GPUs).

    Returns (stats [b, len(STAT_FIELDS)] float32 on every rank, local attack() tuple, clouds [b,3,n] | None)."""
    from . import attack as atk

    rank, _, world = env_world()
    total = int(input_data[0].shape[0] * input_data[0].shape[1])
    rows = shard_rows(total, world, rank)
    n_stats = len(STAT_FIELDS)
    dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    if len(rows) == 0:  # more ranks than rows: nothing to optimise here, but the collective still needs this rank
        local_stats = torch.zeros(0, n_stats, device=dev)
        result, best = None, torch.zeros((0, 3, int(input_data[0].shape[-1] if input_data[0].shape[2] == 3
                                                  else input_data[0].shape[2])), device=dev)
    else:
        result = atk.attack(net, shard_input(input_data, rows), cfg, ref_quirks=ref_quirks,
                            use_cuda_graph=use_cuda_graph, global_batch=total, rows=rows, seed=seed, return_state=True)
        st = result[-1]
        local_stats, best = state_stats(st), st.best_attack
        result = result[:-1]
    stats = gather_stats(local_stats, total)
    clouds = gather_clouds(best, total) if with_clouds else None
    return stats, result, clouds


def barrier():
    if dist.is_initialized():
        dist.barrier()


def max_over_ranks(value, device):
    """device-side max of a python float over ranks (timing rule: max over ranks, never wall clock)."""
    t = torch.tensor([float(value)], device=device, dtype=torch.float64)
    if dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
