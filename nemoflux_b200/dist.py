"""Time-step sharding over the GPUs of one node + the single collective of the path.

The fluxplot.py:51-59 loop carries no state from one time step to the next, so rank r of R
owns a contiguous block of time steps, computes its (nt_r, M) flux series with K2+K3 on its own GPU, and
ONE allgather (NCCL over NVLink on GPUs; gloo in the CPU tests) assembles the (nt, M) series on every rank.
u/v never cross the link; grid, arc lengths, thickness and the K1 weight lists are replicated.
"""
import numpy


def shard_time(nt, world, rank):
    """(t0, n): rank's contiguous block; the first nt % world ranks get one step more (73 over 8 -> 10,9,...,9)"""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f'bad rank {rank} of {world}')
    base, rem = divmod(int(nt), world)
    n = base + (1 if rank < rem else 0)
    t0 = rank * base + min(rank, rem)
    return t0, n


def shard_counts(nt, world):
    return [shard_time(nt, world, r)[1] for r in range(world)]


def allgather_series(local, nt, group=None):
    """local: (nt_r, M) float64 tensor of this rank (cuda for nccl, cpu for gloo) -> (nt, M) on every rank.

    Shards may differ by one step: each rank pads to the largest count, one all_gather moves
    world * max_count * M doubles (81 920 B per rank for ORCA12 / 1024 transects), the padding is dropped."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        if local.shape[0] != nt:
            raise ValueError('single process: the local series must cover every time step')
        return local
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    counts = shard_counts(nt, world)
    if local.shape[0] != counts[rank]:
        raise ValueError(f'rank {rank} must hold {counts[rank]} time steps, got {local.shape[0]}')
    cmax = max(counts)
    m = local.shape[1]
    padded = local
    if local.shape[0] != cmax:
        padded = torch.zeros((cmax, m), dtype=local.dtype, device=local.device)
        padded[:local.shape[0]] = local
    padded = padded.contiguous()
    out = torch.empty((world * cmax, m), dtype=local.dtype, device=local.device)
    if local.is_cuda:
        dist.all_gather_into_tensor(out, padded, group=group)
    else:
        chunks = list(out.view(world, cmax, m).unbind(0))
        dist.all_gather(chunks, padded, group=group)
    if all(c == cmax for c in counts):
        return out
    return torch.cat([out[r * cmax:r * cmax + counts[r]] for r in range(world)], dim=0)


def allreduce_max(value, device=None, group=None):
    """max over ranks of a python float (Field.maxAbsFlux semantics, field.py:234)"""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or 'cpu')
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.cpu()[0])


def sharded_flux_series(compute_local, nt, group=None):
    """compute_local(t0, n) -> (n, M) tensor for time steps t0..t0+n-1; returns the gathered (nt, M)"""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    else:
        world, rank = 1, 0
    t0, n = shard_time(nt, world, rank)
    return allgather_series(compute_local(t0, n), nt, group=group)


def check_partition(nt, world):
    """True when the shards tile 0..nt-1 exactly (used by the tests)"""
    covered = numpy.zeros(nt, int)
    for r in range(world):
        t0, n = shard_time(nt, world, r)
        covered[t0:t0 + n] += 1
    return bool((covered == 1).all())
