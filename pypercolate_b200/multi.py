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


def exchange_comm_id_file(path, rank, timeout=300.0):
    """Hand rank 0's communicator id to the other ranks through a file on a shared file system
    (the transport the reference's Jug workers use for everything): rank 0 writes ``path``
    atomically, the others wait for it.  No torch, no MPI."""
    import os
    import time
    from . import _native
    if rank == 0:
        comm_id = _native.comm_unique_id()
        tmp = "%s.tmp.%d" % (path, os.getpid())
        with open(tmp, "wb") as f:
            f.write(comm_id)
        os.replace(tmp, path)
        return comm_id
    t0 = time.time()
    while True:
        try:
            with open(path, "rb") as f:
                comm_id = f.read()
            if len(comm_id) == _native.COMM_ID_BYTES:
                return comm_id
        except IOError:
            pass
        if time.time() - t0 > timeout:
            raise RuntimeError("no communicator id at %r after %.0f s" % (path, timeout))
        time.sleep(0.01)


def exchange_comm_id_torch(group=None, device=None):
    """The same through an initialised ``torch.distributed`` group (any backend)."""
    import torch
    import torch.distributed as dist
    from . import _native
    rank = dist.get_rank(group)
    raw = _native.comm_unique_id() if rank == 0 else bytes(_native.COMM_ID_BYTES)
    t = torch.tensor(list(raw), dtype=torch.uint8)
    if dist.get_backend(group) == "nccl":
        t = t.to(device if device is not None else torch.device("cuda", torch.cuda.current_device()))
    src = dist.get_global_rank(group, 0) if group is not None else 0
    dist.broadcast(t, src=src, group=group)
    return bytes(t.cpu().tolist())


def ensure_comm(ctx, group=None):
    """Give ``ctx`` a communicator over the ranks of ``group`` (once per context)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if ctx.comm_world != world:
        comm_id = exchange_comm_id_torch(group, torch.device("cuda", ctx.device))
        ctx.comm_init(world, dist.get_rank(group), comm_id)
    return ctx


def allreduce_context(ctx, group=None):
    """Combine the accumulators of every rank's context; afterwards every rank holds the totals.

    One collective step through the C-ABI (``pz_allreduce``: NCCL integer all-reduce of the micro
    accumulators with the run count, all-gather + rank-ordered Chan merge of the canonical
    partials) on the context's own stream.  ``torch.distributed`` only carries the communicator
    id to the ranks the first time.  With a non-NCCL backend (gloo in the CPU tests of the host
    logic) the same exchange is made through host arrays."""
    import torch.distributed as dist
    if dist.get_backend(group) == "nccl":
        ensure_comm(ctx, group).allreduce()
        return ctx.micro_runs
    return _allreduce_context_host(ctx, group)


def _allreduce_context_host(ctx, group=None):
    """``allreduce_context`` for backends that move host memory (gloo)."""
    import torch
    import torch.distributed as dist
    runs = torch.tensor([ctx.micro_runs], dtype=torch.int64)
    dist.all_reduce(runs, op=dist.ReduceOp.SUM, group=group)
    total_runs = int(runs.item())
    if total_runs > 0:
        words = allreduce_words(ctx.micro_export(), group=group)
        ctx.micro_import(words, total_runs)
    if ctx.num_p:
        count, mean, m2 = ctx.canon_export()
        c, mu, s2 = allgather_merge_canon(count, mean, m2, group=group)
        ctx.canon_replace(c, mu, s2)
    return total_runs
