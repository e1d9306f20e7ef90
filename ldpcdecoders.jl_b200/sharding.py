"""Data-parallel sharding of a syndrome batch over ranks (one process per GPU).

Syndromes are independent (no message crosses columns of batchdecode!,
/root/reference/src/decoders/belief_propagation.jl:224-228), so ranks own contiguous column
ranges and the only exchange is one all-reduce of a handful of int64 counters.
Works with any torch.distributed backend (NCCL on the GPUs, gloo in the CPU tests)."""


def shard_bounds(B, world_size, align=32):
    """Contiguous column ranges [lo[r], lo[r+1]) with boundaries that are multiples of `align`
    (bit-packed formats stay word aligned) -- the same split libldpcb200 uses across devices."""
    blocks = (B + align - 1) // align
    lo = [min(B, (blocks * r // world_size) * align) for r in range(world_size)]
    lo.append(B)
    return lo


def shard_range(B, rank, world_size, align=32):
    lo = shard_bounds(B, world_size, align)
    return lo[rank], lo[rank + 1]


def allreduce_counters(counters, group=None):
    """Sum a 1-D int64 tensor of counters over all ranks (in place) and return it."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(counters, op=dist.ReduceOp.SUM, group=group)
    return counters
