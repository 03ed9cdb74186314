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
