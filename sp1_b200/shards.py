"""Shard -> rank placement and the cross-rank timing reduction used by bench.py (and by a multi-GPU host driver).

Reference behaviour: shards of one execution are proven independently (crates/prover/src/worker/... hands every
`ProveShard` task to whichever GPU worker is free; sp1-gpu runs one prover per device), and there is NO exchange between
shard provers: each has its own challenger and commitments.  So the N>1 path is a static round-robin placement of shard
indices plus two control-plane collectives (barrier, MAX of the per-rank elapsed time); nothing here touches trace data.
"""
import torch
import torch.distributed as dist


def shards_of_rank(n_shards, rank, world):
    """round-robin placement: shard i is proven by rank i % world; returns this rank's shard indices in proving order"""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    return list(range(rank, n_shards, world))


def shard_seed(base_seed, shard_index):
    """synthetic-trace seed of a shard: a function of the SHARD index only, so a shard's proof does not depend on placement"""
    return base_seed + 1000 * (shard_index + 1)


def max_over_ranks(value, device="cpu"):
    """MAX-reduce a python float over the process group (identity when not initialised)"""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device="cpu"):
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def aggregate_throughput(units_this_rank, elapsed_ms_this_rank, device="cpu"):
    """whole-job throughput = units proven by all ranks / max-over-ranks elapsed time (units per second)"""
    total = sum_over_ranks(units_this_rank, device)
    ms = max_over_ranks(elapsed_ms_this_rank, device)
    return total / (ms / 1e3), ms


# ---- shard queue + proof gather (BASELINE config 5 / SURVEY.md 8e: N distinct shards proven by W ranks x k contexts, proofs gathered
# to one rank for the recursion tree).  Reference behaviour: the controller hands every ProveShard task to whichever worker is free
# (crates/prover/src/worker/client.rs:29-69, controller/core.rs) and collects the ShardProofs for the compress tree.
import threading


class ShardQueue:
    """the shards placed on this rank (round-robin over ranks), popped by the rank's in-flight prover contexts as they become free"""

    def __init__(self, n_shards, rank, world):
        self._items = shards_of_rank(n_shards, rank, world)
        self._next = 0
        self._lock = threading.Lock()

    def pop(self):
        with self._lock:
            if self._next >= len(self._items):
                return None
            i = self._items[self._next]
            self._next += 1
            return i

    def __len__(self):
        return len(self._items)


def gather_proofs(proofs, dst=0, device="cpu"):
    """proofs: {shard_index: 1-D uint32 numpy array} proven by this rank.  Returns on rank `dst` the dict of ALL ranks' proofs (None on
    the others).  Two collectives: an all_gather of the (index, length) table, then one gather of the padded words (NCCL over NVLink on
    the GPU box, gloo in the CPU tests) - this is the only data-path collective of the whole job."""
    import numpy as np
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return dict(proofs)
    world, rank = dist.get_world_size(), dist.get_rank()
    idx = sorted(proofs)
    meta = torch.tensor([len(idx)] + [v for i in idx for v in (i, int(proofs[i].size))], dtype=torch.int64, device=device)
    n_meta = torch.tensor([meta.numel()], dtype=torch.int64, device=device)
    sizes = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(sizes, n_meta)
    mmax = int(max(int(s.item()) for s in sizes))
    pad = torch.zeros(mmax, dtype=torch.int64, device=device)
    pad[:meta.numel()] = meta
    metas = [torch.zeros(mmax, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(metas, pad)
    tables = []
    for m in metas:
        m = m.cpu().tolist()
        tables.append([(m[1 + 2 * k], m[2 + 2 * k]) for k in range(m[0])])
    wmax = max([sum(ln for _, ln in t) for t in tables] + [1])
    mine = np.zeros(wmax, np.int32)
    off = 0
    for i in idx:
        mine[off:off + proofs[i].size] = proofs[i].view(np.int32)
        off += proofs[i].size
    t = torch.from_numpy(mine).to(device)
    bufs = [torch.zeros(wmax, dtype=torch.int32, device=device) for _ in range(world)] if rank == dst else None
    dist.gather(t, bufs, dst=dst)
    if rank != dst:
        return None
    out = {}
    for r, tab in enumerate(tables):
        words = bufs[r].cpu().numpy().view(np.uint32)
        off = 0
        for i, ln in tab:
            out[i] = words[off:off + ln].copy()
            off += ln
    return out
