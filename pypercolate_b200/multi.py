"""
Runs sharded over the GPUs of one box: one process per GPU
(``torch.distributed``, NCCL over NVLink; gloo in the CPU tests).

The reference treats runs as a map/reduce over seeds with an associative,
commutative reducer (percolate/hpc.py:642-643; percolate/share/jugfile.py:
126-135, 240-244: pickles on a shared file system).  Here every rank sweeps
its own seeds and folds them into its context; one exchange step at the end
combines the ranks:

* micro accumulators (per-n integer sums, 32-bit limbs in 64-bit words) are
  summed word-wise with ONE all-reduce -- exact, order independent;
* canonical partials ``(count, mean, M2)`` are all-gathered and merged in rank
  order with the Chan et al. update (the arithmetic of ``bond_reduce``), so
  every rank holds bit-identical results.

torch is used for the collective plumbing only.
"""

import numpy as np


def shard_bounds(total, rank, world):
    """Contiguous shard ``[lo, hi)`` of ``total`` runs for ``rank``."""
    base, extra = divmod(int(total), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def chan_merge(count_a, mean_a, m2_a, count_b, mean_b, m2_b):
    """Pairwise merge of two ``(count, mean, M2)`` partials (float64 arrays)."""
    if count_b == 0:
        return count_a, mean_a, m2_a
    if count_a == 0:
        return count_b, mean_b, m2_b
    fa, fb = float(count_a), float(count_b)
    n = fa + fb
    delta = mean_b - mean_a
    mean = mean_a + delta * fb / n
    m2 = m2_a + m2_b + delta * delta * fa * fb / n
    return count_a + count_b, mean, m2


def allreduce_words(words, group=None):
    """Word-wise integer sum over ranks of a uint64 numpy array or an int64
    torch tensor (in place for tensors).  Exact because every accumulator word
    holds a 32-bit limb (see include/pz.h)."""
    import torch
    import torch.distributed as dist
    if isinstance(words, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(words).view(np.int64).copy())
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        return t.numpy().view(np.uint64).reshape(words.shape)
    dist.all_reduce(words, op=dist.ReduceOp.SUM, group=group)
    return words


def allgather_merge_canon(count, mean, m2, group=None, device=None):
    """All-gather ``(count, mean, M2)`` and merge in rank order on every rank."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    mean = np.ascontiguousarray(mean, dtype=np.float64)
    m2 = np.ascontiguousarray(m2, dtype=np.float64)
    payload = torch.from_numpy(np.concatenate(
        [np.array([float(count)]), mean.reshape(-1), m2.reshape(-1)]))
    if device is not None:
        payload = payload.to(device)
    gathered = [torch.empty_like(payload) for _ in range(world)]
    dist.all_gather(gathered, payload, group=group)
    c, mu, s2 = 0, None, None
    k = mean.size
    for g in gathered:
        a = g.cpu().numpy()
        c, mu, s2 = chan_merge(c, mu, s2, int(round(a[0])),
                               a[1:1 + k].reshape(mean.shape),
                               a[1 + k:].reshape(mean.shape))
    return c, mu, s2


def allreduce_context(ctx, group=None):
    """Combine the accumulators of every rank's context (device-resident
    exchange over NCCL).  After the call every rank holds the totals."""
    import torch
    import torch.distributed as dist
    dev = torch.device("cuda", ctx.device)
    runs = torch.tensor([ctx.micro_runs], dtype=torch.int64, device=dev)
    dist.all_reduce(runs, op=dist.ReduceOp.SUM, group=group)
    total_runs = int(runs.item())
    if total_runs > 0:
        from . import _native
        words = torch.empty(((ctx.M + 1) * _native.ACC_WORDS,), dtype=torch.int64, device=dev)
        ctx.micro_export(device_ptr=words.data_ptr())
        torch.cuda.current_stream(dev).synchronize()
        dist.all_reduce(words, op=dist.ReduceOp.SUM, group=group)
        torch.cuda.current_stream(dev).synchronize()
        ctx.micro_import(words.data_ptr(), total_runs, is_device=True)
    if ctx.num_p:
        count, mean, m2 = ctx.canon_export()
        c, mu, s2 = allgather_merge_canon(count, mean, m2, group=group, device=dev)
        ctx.canon_replace(c, mu, s2)
    return total_runs
