"""world_size-2 gloo test of the host-side sharding logic (no GPU): shard_rows covers the batch exactly
once, and gather_stats reassembles per-instance rows in global order with a single collective."""
import os
import socket

import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from geoa3_b200 import dist as gd

    gd.init(backend="gloo")
    rows = gd.shard_rows(total, world, rank)
    idx = torch.tensor(list(rows), dtype=torch.float32)
    pc = torch.zeros(len(rows), 3, 8)
    best = pc + idx[:, None, None] * 0.01
    local = gd.pack_stats(success=(idx % 2 == 0), best_loss=idx * 2, best_step=idx + 1, best_attack=best, pc_ori=pc)
    full = gd.gather_stats(local, total)
    mx = gd.max_over_ranks(float(rank + 1), torch.device("cpu"))
    gd.barrier()
    if rank == 0:
        q.put((full.clone(), mx))
    torch.distributed.destroy_process_group()


def test_shard_rows_partition():
    from geoa3_b200 import dist as gd

    for total, world in ((250, 8), (2000, 8), (5, 8), (7, 2), (1, 1)):
        seen = [i for r in range(world) for i in gd.shard_rows(total, world, r)]
        assert seen == list(range(total))
    assert [len(gd.shard_rows(250, 8, r)) for r in range(8)] == [32] * 7 + [26]


def test_gather_stats_world2_gloo():
    total, world = 7, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    full, mx = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert full.shape == (total, 8)
    idx = torch.arange(total, dtype=torch.float32)
    assert torch.equal(full[:, 0], (idx % 2 == 0).float())
    assert torch.equal(full[:, 1], idx * 2) and torch.equal(full[:, 2], idx + 1)
    assert torch.allclose(full[:, 4], idx * 0.01)
    assert mx == 2.0
