"""N>1 host logic on CPU: world_size-2 gloo run of the utterance partition + stats reduction
used by bench.py / multi-GPU decode (no GPU needed)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from parallelwavegan_b200 import sharding


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = sharding.partition(11, rank, world)
    secs, units = sharding.reduce_stats(1.0 + rank, len(mine), dist=dist)
    q.put((rank, mine, secs, units))
    dist.destroy_process_group()


def test_partition_covers_everything_once():
    for n in (0, 1, 7, 16):
        for world in (1, 2, 3, 8):
            seen = sorted(i for r in range(world) for i in sharding.partition(n, r, world))
            assert seen == list(range(n))


def test_two_rank_gloo_reduction():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert res[0][1] == [0, 2, 4, 6, 8, 10] and res[1][1] == [1, 3, 5, 7, 9]
    for _, _, secs, units in res:
        assert secs == 2.0 and units == 11.0  # max over ranks, sum over ranks
