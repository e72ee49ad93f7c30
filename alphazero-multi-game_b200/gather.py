"""Multi-GPU plumbing of the self-play path: games are independent, so slots are sharded over ranks with no
collective during the waves; the only exchange is the gather of finished-game sample records and the sum of the counters.
Backend-agnostic (`nccl` on the GPUs, `gloo` in the CPU tests): everything here is torch.distributed on plain tensors."""
import torch


def shard_slots(total_slots: int, world: int, rank: int):
    """Contiguous block of game slots owned by `rank` (first, count); blocks differ by at most one slot."""
    base, extra = divmod(total_slots, world)
    count = base + (1 if rank < extra else 0)
    first = rank * base + min(rank, extra)
    return first, count


def all_gather_samples(dist, records: torch.Tensor, n_local: int, record_bytes: int):
    """records: uint8 tensor holding `capacity` fixed-size records of which the first n_local are valid (the engine's
    az_engine_drain_samples_device output).  Every rank gets all valid records of all ranks, rank order preserved:
    returns (uint8 tensor [total, record_bytes], list of per-rank counts).  A count vector first, then equal chunks of max(count)
    records (not the whole fixed-capacity buffer), no all-to-all: < 1 MB/s/GPU at the tensor-bound move rate."""
    world = dist.get_world_size()
    cap = records.numel() // record_bytes
    assert 0 <= n_local <= cap
    counts = torch.zeros(world, dtype=torch.int64, device=records.device)
    dist.all_gather_into_tensor(counts, torch.tensor([n_local], dtype=torch.int64, device=records.device))
    cl = [int(c) for c in counts.tolist()]
    m = max(cl)                                     # every rank sends max(count) records: equal-size chunks, but only as many as are needed
    rec = records.view(cap, record_bytes)
    if m == 0:
        return rec[:0], cl
    gathered = torch.empty(world * m * record_bytes, dtype=torch.uint8, device=records.device)
    dist.all_gather_into_tensor(gathered, rec[:m].contiguous().view(-1))
    g = gathered.view(world, m, record_bytes)
    out = torch.cat([g[r, :cl[r]] for r in range(world)], 0)
    return out, cl


def all_reduce_stats(dist, stats: dict, device="cpu"):
    """Sum of the engine counters (az_stats) over ranks."""
    keys = sorted(stats)
    t = torch.tensor([float(stats[k]) for k in keys], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return {k: int(v) for k, v in zip(keys, t.tolist())}
