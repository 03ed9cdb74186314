"""N>1 host logic on CPU: gloo, world_size 2 (the data path has no collective; placement + timing reduction only)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sp1_b200 import shards as S


def test_round_robin_covers_every_shard_once():
    for n in (0, 1, 5, 8, 13):
        for world in (1, 2, 4, 8):
            seen = []
            for r in range(world):
                seen += S.shards_of_rank(n, r, world)
            assert sorted(seen) == list(range(n))
    with pytest.raises(ValueError):
        S.shards_of_rank(4, 2, 2)


def test_seed_depends_on_shard_not_rank():
    assert S.shard_seed(42, 3) == S.shard_seed(42, 3)
    assert len({S.shard_seed(42, i) for i in range(64)}) == 64


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = S.shards_of_rank(5, rank, world)
        # pretend rank r needs 10 ms per shard plus r ms of skew
        elapsed = 10.0 * len(mine) + rank
        units = 100.0 * len(mine)
        thr, ms = S.aggregate_throughput(units, elapsed)
        dist.barrier()
        q.put((rank, mine, thr, ms))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_aggregate():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == [0, 2, 4] and res[1][1] == [1, 3]
    # rank 0: 30 ms, rank 1: 21 ms -> max 30 ms; 500 units in 30 ms
    for _, _, thr, ms in res:
        assert ms == pytest.approx(30.0)
        assert thr == pytest.approx(500.0 / 0.030)
