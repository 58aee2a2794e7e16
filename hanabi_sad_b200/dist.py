"""Multi-GPU plumbing (one process per GPU, torch.distributed): games shard across ranks with NO actor-side collective
(SURVEY.md 8e); the only exchange of the whole system is the learner's gradient all-reduce, plus bookkeeping reductions.

    shard_games(total, rank, world)      -> (first game index, number of games) of this rank
    rank_seed(seed, rank)                -> engine seed of this rank (distinct Philox streams per rank)
    allreduce_gradients(params, world)   -> one flat NCCL/gloo all-reduce (sum) of all grads, averaged
    reduce_throughput(units, ms)         -> whole-job units/s with max-over-ranks time (bench.py contract)
    replay_union(size, weight_sum)       -> (N, sum_w) of the replay shards of all ranks (one SUM all-reduce)
    normalize_importance_weights(raw)    -> raw / max over ALL ranks (one MAX all-reduce)
"""
import torch
import torch.distributed as dist


def shard_games(total_games, rank, world):
    base, rem = divmod(int(total_games), int(world))
    n = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, n


def rank_seed(seed, rank):
    return int(seed) + 1000003 * int(rank)


def allreduce_gradients(params, world=None):
    """Data-parallel learner replicas: flatten every grad into one buffer (18.6 MB fp32 for the R2D2 net), one all-reduce
    over NVLink/NVSwitch, scatter back averaged.  Latency-bound at this size, so a single bucket is the right shape."""
    params = [p for p in params if p.grad is not None]
    if not params:
        return 0
    world = world or (dist.get_world_size() if dist.is_initialized() else 1)
    if world == 1:
        return sum(p.grad.numel() for p in params)
    flat = torch.cat([p.grad.reshape(-1) for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat.div_(world)
    off = 0
    for p in params:
        n = p.grad.numel()
        p.grad.copy_(flat[off:off + n].view_as(p.grad))
        off += n
    return off


def reduce_throughput(units_this_rank, ms_this_rank, device="cpu"):
    """Sum of units over ranks divided by the MAX time over ranks."""
    t = torch.tensor([float(units_this_rank), float(ms_this_rank)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        u, m = t[:1].clone(), t[1:].clone()
        dist.all_reduce(u, op=dist.ReduceOp.SUM)
        dist.all_reduce(m, op=dist.ReduceOp.MAX)
        t = torch.cat([u, m])
    return float(t[0]) / (float(t[1]) * 1e-3), float(t[1])


def replay_union(shard_size, shard_weight_sum, device="cpu"):
    """The replay is sharded over the ranks (each rank samples its own shard).  Importance weights are defined over the UNION
    (rela/prioritized_replay.h:334-339: (N * P(i))^-beta with N = all entries): returns (N, sum_w) summed over ranks."""
    t = torch.tensor([float(shard_size), float(shard_weight_sum)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t[0]), float(t[1])


def shard_sampling_totals(shard_weight_sum, union_size, world=None):
    """What to pass to Engine.sample(total_weight=, total_size=, normalize=False) on every rank: a rank that draws its share of
    the batch from its OWN shard picks entry i with probability w_i / (R * sum_shard) per batch slot -- the importance weight
    must use that probability (not w_i / sum_union, which would only hold if the shards' sums were equal)."""
    world = world or (dist.get_world_size() if dist.is_initialized() else 1)
    return float(world) * float(shard_weight_sum), float(union_size)


def normalize_importance_weights(raw):
    """weights /= weights.max() (prioritized_replay.h:339) with the maximum taken over the sub-batches of ALL ranks."""
    m = raw.detach().max().reshape(1).clone()
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(m, op=dist.ReduceOp.MAX)
    return raw / m
