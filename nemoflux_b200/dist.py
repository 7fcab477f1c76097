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


def batch_bounds(total, world, weights=None):
    """world + 1 boundaries cutting `total` batches into contiguous ranges: equal (+-1) by default, proportional to
    `weights` (each rank's measured throughput) when given -- GPUs of one box differ by a few per cent in sustained
    HBM rate under a shared power budget, and the slowest one sets the time of a max-over-ranks pass"""
    total = int(total)
    if weights is None:
        base, rem = divmod(total, world)
        return [r * base + min(r, rem) for r in range(world)] + [total]
    w = numpy.asarray(weights, dtype=numpy.float64)
    if w.shape != (world,) or not numpy.all(numpy.isfinite(w)) or numpy.any(w <= 0):
        raise ValueError('weights must be `world` positive finite numbers')
    cum = numpy.concatenate([[0.0], numpy.cumsum(w)]) / w.sum()
    b = [int(round(total * c)) for c in cum]
    b[0], b[-1] = 0, total
    for r in range(1, world + 1):           # monotone
        b[r] = max(b[r], b[r - 1])
    return b


def shard_batches(nt, npanels, world, rank, weights=None):
    """Balanced sharding at (time step, panel of cells) granularity.

    The flattened batch space b = t*npanels + q (nt*npanels batches) is cut into `world` contiguous ranges of
    equal length (+-1 batch): 73 snapshots x 26 panels over 8 ranks -> 237 or 238 batches = 9.125 time steps each
    instead of 10,9,...,9 -- or, with `weights`, of lengths proportional to each rank's measured rate.  Returns
    dict(t_first, nt_touched, b0, b1): the rank needs the time steps t_first .. t_first+nt_touched-1 in memory and
    runs the LOCAL batch range [b0, b1) (indices relative to t_first*npanels) with
    PolylineIntegral.fluxSeries(batch_range=(b0, b1))."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f'bad rank {rank} of {world}')
    bounds = batch_bounds(int(nt) * int(npanels), world, weights)
    g0, g1 = bounds[rank], bounds[rank + 1]
    if g1 == g0:
        return dict(t_first=0, nt_touched=0, b0=0, b1=0, g0=g0, g1=g1)
    t_first = g0 // npanels
    t_last = (g1 - 1) // npanels
    return dict(t_first=t_first, nt_touched=t_last - t_first + 1, b0=g0 - t_first * npanels,
                b1=g1 - t_first * npanels, g0=g0, g1=g1)


def combine_partial_series(partial, nt, npanels, group=None, weights=None):
    """partial: this rank's (nt_touched, M) partial sums from fluxSeries(batch_range=...) for its shard_batches
    range -> the full (nt, M) series on every rank.  One all_gather of the (padded) partial blocks; a time step
    shared by several ranks is summed in rank order (deterministic)."""
    import torch
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    else:
        world, rank = 1, 0
    shards = [shard_batches(nt, npanels, world, r, weights) for r in range(world)]
    need = shards[rank]['nt_touched']
    if partial.shape[0] != need:
        raise ValueError(f'rank {rank} must hold {need} time steps, got {partial.shape[0]}')
    m = partial.shape[1]
    cmax = max(max(s['nt_touched'] for s in shards), 1)
    padded = torch.zeros((cmax, m), dtype=partial.dtype, device=partial.device)
    padded[:partial.shape[0]] = partial
    if world > 1:
        gathered = torch.empty((world * cmax, m), dtype=partial.dtype, device=partial.device)
        if partial.is_cuda:
            dist.all_gather_into_tensor(gathered, padded, group=group)
        else:
            dist.all_gather(list(gathered.view(world, cmax, m).unbind(0)), padded, group=group)
    else:
        gathered = padded
    out = torch.zeros((nt, m), dtype=partial.dtype, device=partial.device)
    for r, s in enumerate(shards):                      # rank order -> the same sum on every rank
        n = s['nt_touched']
        if n:
            out[s['t_first']:s['t_first'] + n] += gathered[r * cmax:r * cmax + n]
    return out


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


def gpu_local_cpus(device_index):
    """CPUs NVML reports as local to the GPU (same NUMA node / PCIe root), intersected with the CPUs this process
    may use; an empty set when NVML or the information is not available (containers often hide it)."""
    import os
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        try:
            uuid = str(torch.cuda.get_device_properties(device_index).uuid)
            h = None
            for i in range(pynvml.nvmlDeviceGetCount()):     # CUDA_VISIBLE_DEVICES renumbers: match by UUID
                hi = pynvml.nvmlDeviceGetHandleByIndex(i)
                u = pynvml.nvmlDeviceGetUUID(hi)
                u = u.decode() if isinstance(u, bytes) else u
                if uuid in u:
                    h = hi
                    break
            if h is None:
                return set()
            ncpu = os.cpu_count() or 1
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        finally:
            pynvml.nvmlShutdown()
    except Exception:
        return set()
    cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
    return cpus & set(os.sched_getaffinity(0))


def bind_to_gpu_cpus(device_index):
    """Run this process (and the pinned host buffers it allocates from now on: first touch) on the CPUs next to its
    GPU.  With one process per GPU on a two-socket box this keeps every rank's host->device stream on its own
    socket's memory controllers and PCIe root instead of crossing the socket link.  Returns the number of CPUs the
    process was bound to, 0 if nothing was changed."""
    import os
    cpus = gpu_local_cpus(device_index)
    if not cpus or cpus == set(os.sched_getaffinity(0)):
        return 0
    try:
        os.sched_setaffinity(0, cpus)
    except OSError:
        return 0
    return len(cpus)


def apportion_by_rate(total, rates):
    """Split `total` time steps over the ranks in proportion to `rates` (each rank's measured host->device rate),
    largest-remainder rounding, so that the slowest link no longer sets the time of a host-fed pass: with equal
    shards the rank behind a 23 GB/s link finishes 50 % after the one behind a 35 GB/s link
    (profiles/r1_numa_probe8.md).  Returns the list of counts (sum == total, every count >= 0); a time step is the
    unit, and the series of a time step does not depend on the rank that computes it (fluxplot.py:51-59 carries
    no state), so the gathered series is bit-identical to the equal-shard one."""
    total = int(total)
    r = numpy.asarray(rates, dtype=numpy.float64)
    if r.ndim != 1 or r.size == 0 or total < 0:
        raise ValueError('apportion_by_rate needs a non-empty 1-D list of rates and total >= 0')
    if not numpy.all(numpy.isfinite(r)) or numpy.any(r < 0):
        raise ValueError('rates must be finite and >= 0')
    if r.sum() <= 0:
        r = numpy.ones_like(r)
    ideal = total * r / r.sum()
    counts = numpy.floor(ideal).astype(numpy.int64)
    left = total - int(counts.sum())
    order = numpy.argsort(-(ideal - counts), kind='stable')     # ties: lower rank first
    counts[order[:left]] += 1
    # exchange steps while it lowers the makespan max(count/rate): floor+remainder is optimal for the sum, not the max
    span = lambda c: float(numpy.max(numpy.where(c > 0, c / numpy.maximum(r, 1e-300), 0.0)))   # noqa: E731
    for _ in range(4 * r.size):
        worst = int(numpy.argmax(numpy.where(counts > 0, counts / numpy.maximum(r, 1e-300), -1.0)))
        trial_best, best = None, span(counts)
        for j in range(r.size):
            if j == worst or r[j] <= 0:
                continue
            c = counts.copy()
            c[worst] -= 1
            c[j] += 1
            s = span(c)
            if s < best - 1e-15:
                trial_best, best = c, s
        if trial_best is None:
            break
        counts = trial_best
    return [int(c) for c in counts]


def blocks_from_counts(counts):
    """[(t0, n)] contiguous blocks of time steps for the per-rank counts"""
    out, t0 = [], 0
    for c in counts:
        out.append((t0, int(c)))
        t0 += int(c)
    return out


def measure_h2d_rate(device, nbytes=256 << 20, reps=3, group=None):
    """This rank's host->device rate in GB/s while EVERY rank copies at the same time (pinned buffer, CUDA events):
    the figure that matters for a host-fed pass on a box whose links share host memory bandwidth."""
    import torch
    import torch.distributed as dist
    host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    devb = torch.empty(nbytes, dtype=torch.uint8, device=device)
    devb.copy_(host, non_blocking=True)
    torch.cuda.synchronize(device)
    if dist.is_available() and dist.is_initialized():
        dist.barrier(group=group)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        devb.copy_(host, non_blocking=True)
    e1.record()
    torch.cuda.synchronize(device)
    return reps * nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9
