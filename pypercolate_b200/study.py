"""
Finite-size-scaling study driver: the production caller of the hpc path.

The reference ships this as a Jug task graph, ``percolate/share/jugfile.py``:
for every system size L it prepares the percolation graph (``:42-47``), the
binomial weights of every p (``:50-54``), draws ``NUMBER_OF_RUNS`` seeds from
one master ``RandomState`` (``:172, 215``), maps ``bond_run`` over the seeds
in ``NUMBER_OF_TASKS`` tasks and folds the per-run canonical statistics with
``bond_reduce`` (``:57-135, 217-244``), finalises (``:247-253``) and appends
one HDF5 dataset keyed by L (``:138-156``).

Here one call does the same per L on the GPU(s): seeds -> fused device pass
(``hpc.bond_canonical_averages_batch``) -> ``finalize_canonical_averages``.
There is no task scheduler to replace: the runs of one size are one batch of
the sweep kernel, sharded over the ranks of ``torch.distributed`` when it is
initialised.  With ``rng='mt19937'`` (default) every run uses numpy's
``RandomState(seed).permutation`` stream, so the result is the reference's up
to the association of the floating-point reduction.
"""

import os

import numpy as np
import scipy.stats

from . import hpc, lowering

UINT32_MAX = 4294967296                               # jugfile.py:36
DEFAULT_SEED = 201508061904 % UINT32_MAX              # jugfile.py:37
ALPHA_1SIGMA = 2 * scipy.stats.norm.cdf(-1.0)         # jugfile.py:38


def study_seeds(system_dimensions, number_of_runs, seed=DEFAULT_SEED):
    """Seeds of every system size, drawn like the jugfile does: one master
    ``RandomState(seed)``, ``randint(UINT32_MAX, size=number_of_runs)`` per
    dimension in the order given (jugfile.py:172, 215)."""
    rng = np.random.RandomState(seed=seed)
    return {int(L): rng.randint(UINT32_MAX, size=int(number_of_runs)).astype(np.uint32)
            for L in system_dimensions}


def _plain_names(array):
    """Same memory layout, field names as plain ``str`` (the reference builds its
    dtypes with ``np.str_`` keys, percolate/hpc.py:22-31, which NumPy 2 cannot
    round-trip through the .npy header)."""
    array = np.ascontiguousarray(array)
    dt = array.dtype
    if dt.names is None:
        return array
    plain = np.dtype({'names': [str(n) for n in dt.names],
                      'formats': [dt.fields[n][0] for n in dt.names],
                      'offsets': [dt.fields[n][1] for n in dt.names],
                      'itemsize': dt.itemsize})
    return array.view(plain)


def write_to_disk(path, dimension, canonical_averages):
    """Append the finalised averages of one size (jugfile.py:138-156): an HDF5
    dataset named ``str(dimension)`` when h5py is importable and ``path`` ends
    in .h5/.hdf5, else an ``.npz`` archive with the same key.  Like the
    reference, an existing key is an error."""
    key = '{}'.format(dimension)
    canonical_averages = _plain_names(canonical_averages)
    if path.endswith(('.h5', '.hdf5')):
        try:
            import h5py
        except ImportError:
            raise RuntimeError("h5py is not installed: use an .npz path")
        with h5py.File(path, mode='a') as f:
            if key in f:
                raise RuntimeError("dataset %r exists already" % key)
            f.create_dataset(name=key, data=canonical_averages)
        return
    existing = {}
    if os.path.exists(path):
        with np.load(path) as z:
            existing = {k: z[k] for k in z.files}
    if key in existing:
        raise RuntimeError("dataset %r exists already" % key)
    existing[key] = canonical_averages
    tmp = path + '.tmp.npz'
    np.savez(tmp, **existing)
    os.replace(tmp, path)


def finite_size_study(system_dimensions=(8, 16, 32), number_of_runs=10000,
                      ps=None, alpha=ALPHA_1SIGMA, seed=DEFAULT_SEED,
                      spanning_cluster=True, rng='mt19937', output=None, device=None,
                      graph=lowering.lowered_spanning_2d_grid):
    """The jugfile's study in one call.

    Returns ``{L: finalized canonical averages}`` (dtype
    ``hpc.finalized_canonical_averages_dtype``), one row per p.  ``graph`` maps
    a size to a lowered graph (default: the closed-form ``spanning_2d_grid``;
    any callable returning a networkx graph with spanning sides, a
    ``percolation_graph`` dict or a ``LoweredGraph`` works).  ``output``: file
    to append each size to (see ``write_to_disk``).  Under an initialised
    ``torch.distributed`` every rank passes the same arguments, sweeps its
    shard of the seeds and receives the full result; rank 0 writes.
    """
    ps = np.linspace(0.4, 0.6, num=40) if ps is None else np.asarray(ps, dtype=np.float64)
    seeds = study_seeds(system_dimensions, number_of_runs, seed)
    rank, world = 0, 1
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            rank, world = dist.get_rank(), dist.get_world_size()
    except ImportError:
        pass
    results = {}
    for L in system_dimensions:
        L = int(L)
        g = graph(L)
        if isinstance(g, lowering.LoweredGraph):
            kwargs = dict(perc_graph=g, num_nodes=g.num_nodes, num_edges=g.num_edges)
        else:
            if not isinstance(g, dict):          # a networkx graph with auxiliary nodes
                from . import percolate as _percolate
                g = _percolate.percolation_graph(g, spanning_cluster=spanning_cluster)
            # a percolation_graph() dict carries its own 'graph' and 'spanning_cluster' entries
            # (percolate/percolate.py:55-100); the study's spanning_cluster argument decides
            kwargs = {k: v for k, v in g.items() if k not in ('graph', 'spanning_cluster')}
        lowered_nodes = kwargs['num_nodes']
        my = seeds[L]
        if world > 1:
            from . import multi
            lo, hi = multi.shard_bounds(my.size, rank, world)
            my = my[lo:hi]
        averages = hpc.bond_canonical_averages_batch(
            seeds=my, ps=ps, spanning_cluster=spanning_cluster, rng=rng, device=device,
            distributed=world > 1, **kwargs)
        final = hpc.finalize_canonical_averages(
            number_of_nodes=lowered_nodes, ps=ps, canonical_averages=averages, alpha=alpha)
        results[L] = final
        if output is not None and rank == 0:
            write_to_disk(output, L, final)
    return results
