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


def test_shard_queue_hands_out_each_local_shard_once():
    import threading
    q = S.ShardQueue(13, rank=1, world=4)
    got, lock = [], threading.Lock()

    def work():
        while True:
            i = q.pop()
            if i is None:
                return
            with lock:
                got.append(i)
    ths = [threading.Thread(target=work) for _ in range(3)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    assert sorted(got) == S.shards_of_rank(13, 1, 4) and len(q) == len(got)


def _gather_worker(rank, world, port, q):
    import numpy as np
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sq = S.ShardQueue(5, rank, world)
        proofs = {}
        while True:
            i = sq.pop()
            if i is None:
                break
            proofs[i] = (np.arange(100 + 37 * i, dtype=np.uint32) * np.uint32(2654435761)) ^ np.uint32(i)   # ragged lengths
        allp = S.gather_proofs(proofs, dst=0)
        dist.barrier()
        q.put((rank, None if allp is None else {k: (int(v.size), int(v.sum())) for k, v in allp.items()}))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_proof_gather():
    """world 2: every shard's proof words arrive on rank 0 exactly once, ragged lengths intact; other ranks get nothing"""
    import numpy as np
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gather_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[1] is None
    want = {}
    for i in range(5):
        v = (np.arange(100 + 37 * i, dtype=np.uint32) * np.uint32(2654435761)) ^ np.uint32(i)
        want[i] = (int(v.size), int(v.sum()))
    assert res[0] == want
