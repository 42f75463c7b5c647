# encoding: utf-8

"""
Low-level routines of the Newman-Ziff algorithm -- B200-native drop-in for
``percolate.hpc`` of andsor/pypercolate (reference: percolate/hpc.py).

Same function names, argument meaning, structured-array fields and error
behaviour as the reference.  The per-bond work runs in hand-written sm_100a
CUDA kernels behind the C-ABI of ``include/pz.h``; there is no CPU fallback
(a missing extension or GPU raises).

Additions (not in the reference) are the batch entry points at the end:
``bond_microcanonical_statistics_batch`` and ``bond_canonical_averages_batch``
-- the device-side form of the reference's map/reduce over seeds
(percolate/share/jugfile.py:57-135).

See also
--------

percolate : The high-level module
"""

import numpy as np
import scipy.stats

from . import _native
from . import lowering as _lowering


# Field tables of the four structured types.  The third entry tells whether the field
# exists only when a spanning cluster is detected.
_MICRO_FIELDS = (
    ('n', 'uint32', False), ('edge', 'uint32', False),
    ('has_spanning_cluster', 'bool', True),
    ('max_cluster_size', 'uint32', False), ('moments', '(5,)uint64', False),
)
_CANON_FIELDS = (
    ('percolation_probability', 'float64', True),
    ('max_cluster_size', 'float64', False), ('moments', '(5,)float64', False),
)
_AVERAGE_FIELDS = (
    ('number_of_runs', 'uint32', False),
    ('percolation_probability_mean', 'float64', True),
    ('percolation_probability_m2', 'float64', True),
    ('max_cluster_size_mean', 'float64', False), ('max_cluster_size_m2', 'float64', False),
    ('moments_mean', '(5,)float64', False), ('moments_m2', '(5,)float64', False),
)
_FINAL_FIELDS = (
    ('number_of_runs', 'uint32', False), ('p', 'float64', False), ('alpha', 'float64', False),
    ('percolation_probability_mean', 'float64', True),
    ('percolation_probability_std', 'float64', True),
    ('percolation_probability_ci', '(2,)float64', True),
    ('percolation_strength_mean', 'float64', False),
    ('percolation_strength_std', 'float64', False),
    ('percolation_strength_ci', '(2,)float64', False),
    ('moments_mean', '(5,)float64', False), ('moments_std', '(5,)float64', False),
    ('moments_ci', '(5,2)float64', False),
)


def _ndarray_dtype(fields, spanning_cluster=True):
    """Packed NumPy structured dtype spec (a list of ``(name, type)``) from a
    field table; same spelling as the reference's helper percolate/hpc.py:22-31."""
    return [(np.str_(name), kind) for name, kind, span_only in fields
            if spanning_cluster or not span_only]


def microcanonical_statistics_dtype(spanning_cluster=True):
    """
    Return the numpy structured array data type for sample states

    Reference: percolate/hpc.py:34-70.  Packed, 53 bytes per row (52 without
    the spanning flag): ``n:u4, edge:u4, [has_spanning_cluster:?],
    max_cluster_size:u4, moments:(5,)u8``.
    """
    return _ndarray_dtype(_MICRO_FIELDS, spanning_cluster)


def _default_device():
    import os
    return int(os.environ.get("PZ_DEVICE", os.environ.get("LOCAL_RANK", "0")))


def _lower(perc_graph, spanning_cluster, auxiliary_node_attributes,
           auxiliary_edge_attributes, spanning_sides):
    return _lowering.lower(
        perc_graph, spanning_cluster=spanning_cluster,
        auxiliary_node_attributes=auxiliary_node_attributes,
        auxiliary_edge_attributes=auxiliary_edge_attributes,
        spanning_sides=spanning_sides,
    )


def _is_u32_seed(seed):
    try:
        import operator
        s = operator.index(seed)
    except TypeError:
        return False
    return 0 <= s <= 0xFFFFFFFF


def _run_rows(lowered, seeds, device=None, rng='mt19937'):
    """Rows of the runs seeded by ``seeds`` (shape (R, M+1))."""
    ctx = _native.context_for(lowered, _default_device() if device is None else device)
    seeds = list(seeds)
    if rng not in _native.RNG_MODES:
        raise ValueError("rng must be one of %s" % sorted(_native.RNG_MODES))
    if rng != 'mt19937':
        return ctx.run_rows(len(seeds), _native.RNG_MODES[rng],
                            np.asarray(seeds, dtype=np.uint32))
    if all(_is_u32_seed(s) for s in seeds):
        # numpy's legacy stream reproduced on the device, bit for bit
        return ctx.run_rows(len(seeds), _native.PERM_MT19937,
                            np.asarray(seeds, dtype=np.uint32))
    # None / array_like seeds: draw on the host exactly like the reference
    # (percolate/hpc.py:195,206)
    perms = np.empty((len(seeds), lowered.num_edges), dtype=np.int32)
    for r, s in enumerate(seeds):
        perms[r] = np.random.RandomState(seed=s).permutation(lowered.num_edges)
    return ctx.run_rows(len(seeds), _native.PERM_HOST, perms)


def bond_sample_states(
    perc_graph, num_nodes, num_edges, seed, spanning_cluster=True,
    auxiliary_node_attributes=None, auxiliary_edge_attributes=None,
    spanning_sides=None,
    **kwargs
):
    '''
    Generate successive sample states of the bond percolation model

    Drop-in for percolate/hpc.py:73-307.  This is a generator function; the
    whole run is computed on the GPU when the generator is first advanced
    (errors surface on the first ``next()`` like in the reference), and the
    states are then handed out one at a time.
    CAUTION: like the reference it returns a reference to one internal 1-row
    array, not a copy (hpc.py:85,221,307).

    Parameters are those of the reference.  ``perc_graph`` may also be a
    ``pypercolate_b200.lowering.LoweredGraph``.  Backend-only keyword
    arguments: ``device`` (CUDA device index), ``rng`` (``'mt19937'`` --
    numpy's ``RandomState(seed).permutation`` reproduced bit for bit, the
    default --, ``'philox'`` (Philox4x32-10 bucketed Fisher-Yates) or
    ``'feistel'`` (Philox-keyed Feistel bijection, the throughput mode)).

    Yields
    ------
    ret : ndarray
        Structured array with dtype ``microcanonical_statistics_dtype``:
        ``n``, ``edge`` (undefined for n == 0), ``has_spanning_cluster`` (only
        if ``spanning_cluster``), ``max_cluster_size``, ``moments`` (k = 0..4,
        uint64, without one largest cluster).

    Raises
    ------
    ValueError
        ``spanning_cluster`` is True but ``spanning_sides`` does not hold
        exactly two sides.
    '''
    if spanning_cluster and not isinstance(perc_graph, _lowering.LoweredGraph):
        if len(spanning_sides) != 2:
            raise ValueError(
                'Spanning cluster is to be detected, but auxiliary nodes '
                'of less or more than 2 types (sides) given.'
            )

    lowered = _lower(perc_graph, spanning_cluster, auxiliary_node_attributes,
                     auxiliary_edge_attributes, spanning_sides)
    if lowered.num_nodes != num_nodes or lowered.num_edges != num_edges:
        raise ValueError('num_nodes / num_edges do not match perc_graph')

    rows = _run_rows(lowered, [seed], kwargs.get('device'),
                     kwargs.get('rng', 'mt19937'))[0]

    ret = np.empty(
        1, dtype=microcanonical_statistics_dtype(spanning_cluster)
    )
    for n in range(num_edges + 1):
        ret[0] = rows[n]
        yield ret


def bond_microcanonical_statistics(
    perc_graph, num_nodes, num_edges, seed,
    spanning_cluster=True,
    auxiliary_node_attributes=None, auxiliary_edge_attributes=None,
    spanning_sides=None,
    **kwargs
):
    """
    Evolve a single run over all microstates (bond occupation numbers)

    Drop-in for percolate/hpc.py:310-404: the structured array of all
    ``num_edges + 1`` states of one run, fields as in ``bond_sample_states``.
    """
    lowered = _lower(perc_graph, spanning_cluster, auxiliary_node_attributes,
                     auxiliary_edge_attributes, spanning_sides)
    if lowered.num_nodes != num_nodes or lowered.num_edges != num_edges:
        raise ValueError('num_nodes / num_edges do not match perc_graph')
    rows = _run_rows(lowered, [seed], kwargs.get('device'),
                     kwargs.get('rng', 'mt19937'))[0]
    return np.ascontiguousarray(rows).astype(
        np.dtype(microcanonical_statistics_dtype(spanning_cluster)), copy=False)


def canonical_statistics_dtype(spanning_cluster=True):
    """
    The NumPy Structured Array type for canonical statistics

    Reference: percolate/hpc.py:407-440.
    """
    return _ndarray_dtype(_CANON_FIELDS, spanning_cluster)


_util_ctx = {}


def _utility_context(device=None):
    """Graph-less context for the calls whose arguments are arrays only."""
    device = _default_device() if device is None else device
    ctx = _util_ctx.get(device)
    if ctx is None:
        ctx = _native.Context(device)
        _util_ctx[device] = ctx
    return ctx


def bond_canonical_statistics(
    microcanonical_statistics,
    convolution_factors,
    **kwargs
):
    """
    canonical cluster statistics for a single run and a single probability

    Drop-in for percolate/hpc.py:443-515: ``sum_n f[n] * Q[n]`` for the
    spanning flag, the largest cluster and the five moments of one
    materialised run; the contraction runs on the GPU.
    """
    spanning_cluster = (
        'has_spanning_cluster' in microcanonical_statistics.dtype.names
    )
    rows = np.ascontiguousarray(
        microcanonical_statistics,
        dtype=np.dtype(microcanonical_statistics_dtype(spanning_cluster)))
    f = np.ascontiguousarray(convolution_factors, dtype=np.float64)
    if f.shape != rows.shape:
        raise ValueError('convolution_factors must have one entry per state')
    out = _utility_context(kwargs.get('device')).canonical_statistics_rows(
        rows, f, spanning_cluster)

    ret = np.empty(1, dtype=canonical_statistics_dtype(spanning_cluster))
    if spanning_cluster:
        ret['percolation_probability'] = out[0]
    ret['max_cluster_size'] = out[1]
    ret['moments'] = out[2:7]
    return ret


def canonical_averages_dtype(spanning_cluster=True):
    """
    The NumPy Structured Array type for canonical averages over several
    runs

    Reference: percolate/hpc.py:518-558.
    """
    return _ndarray_dtype(_AVERAGE_FIELDS, spanning_cluster)


def bond_initialize_canonical_averages(
    canonical_statistics, **kwargs
):
    """
    Initialize the canonical averages from a single-run cluster statistics

    Drop-in for percolate/hpc.py:561-635 (``number_of_runs = 1``, mean = the
    run's value, M2 = 0).  num_p rows of 15 doubles: host arithmetic.
    """
    names = canonical_statistics.dtype.names
    out = np.empty_like(
        canonical_statistics,
        dtype=canonical_averages_dtype('percolation_probability' in names))
    out['number_of_runs'] = 1
    for name in names:                      # one run: mean = its value, no spread yet
        out[name + '_mean'] = canonical_statistics[name]
        out[name + '_m2'] = 0.0
    return out


def _online_variance(n_a, mean_a, m2_a, n_b, mean_b, m2_b):
    """Pairwise merge of ``(n, mean, M2)`` (Chan et al.), the arithmetic the
    reference delegates to ``simoa.stats.online_variance``
    (percolate/hpc.py:677-684; docs/pypercolate-hpc.rst:66-68).  ``n`` is
    converted to float64 before forming ``n_a * n_b`` (it is uint32 in the
    structured array)."""
    n_a = np.asarray(n_a, dtype=np.float64)
    n_b = np.asarray(n_b, dtype=np.float64)
    n = n_a + n_b
    delta = mean_b - mean_a
    mean = mean_a + delta * n_b / n
    m2 = m2_a + m2_b + delta * delta * n_a * n_b / n
    return mean, m2


def bond_reduce(row_a, row_b):
    """
    Reduce the canonical averages over several runs

    Drop-in for percolate/hpc.py:638-702: associative and commutative merge of
    two ``canonical_averages_dtype`` arrays.
    """
    out = np.empty_like(row_a)
    stats = ['max_cluster_size', 'moments']
    if all(key in row.dtype.names for row in (row_a, row_b)
           for key in ('percolation_probability_mean', 'percolation_probability_m2')):
        stats.insert(0, 'percolation_probability')
    n_a, n_b = row_a['number_of_runs'], row_b['number_of_runs']
    for stat in stats:
        mean_a, m2_a = row_a[stat + '_mean'], row_a[stat + '_m2']
        # per-p run counts against (num_p,) or (num_p, 5) statistics
        shape = n_a.shape + (1,) * (mean_a.ndim - n_a.ndim)
        out[stat + '_mean'], out[stat + '_m2'] = _online_variance(
            n_a.reshape(shape), mean_a, m2_a,
            n_b.reshape(shape), row_b[stat + '_mean'], row_b[stat + '_m2'])
    out['number_of_runs'] = n_a + n_b
    return out


def finalized_canonical_averages_dtype(spanning_cluster=True):
    """
    The NumPy Structured Array type for finalized canonical averages over
    several runs

    Reference: percolate/hpc.py:705-749.
    """
    return _ndarray_dtype(_FINAL_FIELDS, spanning_cluster)


def finalize_canonical_averages(
    number_of_nodes, ps, canonical_averages, alpha,
):
    """
    Finalize canonical averages

    Drop-in for percolate/hpc.py:752-834: sample mean, sample standard
    deviation ``sqrt(M2 / (n - 1))`` and Student-t confidence interval; the
    largest cluster and the moments are divided by the number of nodes, the
    percolation probability is not.  scipy is called with the reference's
    arguments, so the quantiles are the reference's by construction.
    """
    names = canonical_averages.dtype.names
    spanning_cluster = ('percolation_probability_mean' in names and
                        'percolation_probability_m2' in names)
    out = np.empty_like(
        canonical_averages,
        dtype=finalized_canonical_averages_dtype(spanning_cluster))
    runs = canonical_averages['number_of_runs']
    out['number_of_runs'] = runs
    out['p'] = ps
    out['alpha'] = alpha

    # (input statistic, output statistic, reported per node?)
    plan = [('max_cluster_size', 'percolation_strength', True), ('moments', 'moments', True)]
    if spanning_cluster:
        plan.insert(0, ('percolation_probability', 'percolation_probability', False))
    for src, dst, per_node in plan:
        mean = out[dst + '_mean']
        std = out[dst + '_std']
        # run counts broadcast against (num_p,) or (num_p, 5) statistics
        n = runs.reshape(runs.shape + (1,) * (mean.ndim - runs.ndim))
        mean[...] = canonical_averages[src + '_mean']
        with np.errstate(divide='ignore', invalid='ignore'):
            std[...] = np.sqrt(canonical_averages[src + '_m2'] / (n - 1))
        if per_node:
            mean /= number_of_nodes
            std /= number_of_nodes
        with np.errstate(divide='ignore', invalid='ignore'):
            lo, hi = scipy.stats.t.interval(1 - alpha, df=n - 1, loc=mean,
                                            scale=std / np.sqrt(n))
        out[dst + '_ci'][..., 0] = lo
        out[dst + '_ci'][..., 1] = hi
    return out


# ---------------------------------------------------------------------------
# batch entry points (beyond the reference surface)
# ---------------------------------------------------------------------------

def bond_microcanonical_statistics_batch(
    perc_graph, num_nodes, num_edges, seeds, spanning_cluster=True,
    auxiliary_node_attributes=None, auxiliary_edge_attributes=None,
    spanning_sides=None, **kwargs
):
    """``bond_microcanonical_statistics`` for many seeds in one device batch:
    array of shape ``(len(seeds), num_edges + 1)``."""
    lowered = _lower(perc_graph, spanning_cluster, auxiliary_node_attributes,
                     auxiliary_edge_attributes, spanning_sides)
    rows = _run_rows(lowered, seeds, kwargs.get('device'),
                     kwargs.get('rng', 'mt19937'))
    return rows.astype(
        np.dtype(microcanonical_statistics_dtype(spanning_cluster)), copy=False)


def _canonical_averages_from_partials(count, mean, m2, spanning_cluster):
    ret = np.empty(mean.shape[0],
                   dtype=canonical_averages_dtype(spanning_cluster))
    ret['number_of_runs'] = count
    if spanning_cluster:
        ret['percolation_probability_mean'] = mean[:, 0]
        ret['percolation_probability_m2'] = m2[:, 0]
    ret['max_cluster_size_mean'] = mean[:, 1]
    ret['max_cluster_size_m2'] = m2[:, 1]
    ret['moments_mean'] = mean[:, 2:7]
    ret['moments_m2'] = m2[:, 2:7]
    return ret


def bond_canonical_averages_batch(
    perc_graph, num_nodes, num_edges, seeds, ps, spanning_cluster=True,
    auxiliary_node_attributes=None, auxiliary_edge_attributes=None,
    spanning_sides=None, **kwargs
):
    """Device-side ``reduce(bond_reduce, map(bond_run, seeds))``
    (percolate/share/jugfile.py:57-135): every run is swept, convolved with the
    binomial weights of every ``p`` (hpc.py:488-515) and folded into
    ``(number_of_runs, mean, M2)`` (hpc.py:607-702) without leaving the GPU.
    Returns an array of ``canonical_averages_dtype`` ready for
    ``finalize_canonical_averages`` or further ``bond_reduce``.
    Backend keyword arguments: ``rng`` (``'mt19937'`` default, ``'philox'``,
    ``'feistel'``), ``device``, ``distributed`` (``seeds`` is this rank's shard;
    combine the partials of all ranks of ``torch.distributed``)."""
    lowered = _lower(perc_graph, spanning_cluster, auxiliary_node_attributes,
                     auxiliary_edge_attributes, spanning_sides)
    device = kwargs.get('device')
    ctx = _native.context_for(lowered, _default_device() if device is None else device)
    rng = kwargs.get('rng', 'mt19937')
    seeds = np.asarray(seeds, dtype=np.uint32)
    ctx.set_ps(np.asarray(ps, dtype=np.float64))
    ctx.reset_accumulators()
    if rng not in _native.RNG_MODES:
        raise ValueError("rng must be one of %s" % sorted(_native.RNG_MODES))
    ctx.run_fused(seeds.size, _native.RNG_MODES[rng], seeds, _native.FUSE_CANON)
    if kwargs.get('distributed'):
        # ``seeds`` is this rank's shard: fold the partials of all ranks (rank order)
        from . import multi
        multi.allreduce_context(ctx)
    count, mean, m2 = ctx.canon_export()
    return _canonical_averages_from_partials(count, mean, m2, spanning_cluster)


def bond_statistics_batch(
    perc_graph, num_nodes, num_edges, seeds, ps, alpha,
    spanning_cluster=True, auxiliary_node_attributes=None,
    auxiliary_edge_attributes=None, spanning_sides=None, **kwargs
):
    """One fused device pass over many runs: microcanonical AND canonical
    averaging (BASELINE config 3).

    Every run (one seed each) is swept once on the GPU; its per-bond statistics
    are folded into (a) exact per-n sums over runs, from which the per-n means
    and confidence intervals of ``microcanonical_averages_arrays``
    (percolate/percolate.py:450-705, 968-1064) follow, and (b) per-run
    canonical values for every ``p`` reduced to ``(count, mean, M2)`` and
    finalised like ``finalize_canonical_averages`` (percolate/hpc.py:443-834).
    With ``torch.distributed`` initialised (one process per GPU) the seeds are
    the rank's shard and the partial results are combined with one NCCL
    exchange (``pypercolate_b200.multi``).

    Returns a dict with ``'microcanonical_averages_arrays'`` (same keys and
    normalisation as the reference's), ``'canonical_averages'``
    (``canonical_averages_dtype``), ``'finalized_canonical_averages'``
    (``finalized_canonical_averages_dtype``) and ``'number_of_runs'``.

    Backend keyword arguments: ``rng`` (``'philox'`` default; ``'feistel'``, the
    keyed-bijection throughput mode; or ``'mt19937'`` for numpy's stream bit for bit), ``device``, ``distributed`` (combine
    across ranks; default: whether ``torch.distributed`` is initialised).
    """
    from . import percolate as _percolate
    lowered = _lower(perc_graph, spanning_cluster, auxiliary_node_attributes,
                     auxiliary_edge_attributes, spanning_sides)
    device = kwargs.get('device')
    ctx = _native.context_for(lowered, _default_device() if device is None else device)
    rng = kwargs.get('rng', 'philox')
    if rng not in _native.RNG_MODES:
        raise ValueError("rng must be one of %s" % sorted(_native.RNG_MODES))
    mode = _native.RNG_MODES[rng]
    seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
    ps = np.ascontiguousarray(ps, dtype=np.float64)

    ctx.set_ps(ps)
    ctx.reset_accumulators()
    ctx.run_fused(seeds.size, mode, seeds, _native.FUSE_MICRO | _native.FUSE_CANON)

    distributed = kwargs.get('distributed')
    if distributed is None:
        try:
            import torch.distributed as dist
            distributed = dist.is_available() and dist.is_initialized()
        except ImportError:
            distributed = False
    if distributed:
        from . import multi
        multi.allreduce_context(ctx)

    runs = ctx.micro_runs
    # means, Student-t intervals and the per-site normalisation are evaluated on the
    # device (pz_micro_arrays: the host formulas, operation for operation)
    arrays = _percolate._arrays_on_device(ctx, runs, alpha, lowered.num_nodes,
                                          lowered.num_edges, spanning_cluster,
                                          norm=lowered.num_nodes)
    count, cmean, cm2 = ctx.canon_export()
    averages = _canonical_averages_from_partials(count, cmean, cm2, spanning_cluster)
    finalized = finalize_canonical_averages(lowered.num_nodes, ps, averages, alpha)
    ctx.reset_accumulators()
    return {
        'microcanonical_averages_arrays': arrays,
        'canonical_averages': averages,
        'finalized_canonical_averages': finalized,
        'number_of_runs': runs,
    }
