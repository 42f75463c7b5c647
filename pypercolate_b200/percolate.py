# encoding: utf-8

"""
Newman-Ziff bond percolation, high-level module -- B200-native drop-in for
``percolate.percolate`` of andsor/pypercolate (reference:
percolate/percolate.py).

Names, signatures, dictionary keys and error behaviour are the reference's.
The per-bond sweep, the reduction over runs and the binomial convolution run
in sm_100a CUDA kernels behind ``include/pz.h``; the scipy quantile calls of
the reference (Student-t and beta) stay on the host and receive the
reference's arguments.  There is no CPU fallback.
"""

import copy

import numpy as np
import scipy.stats
import networkx as nx

from . import _native
from . import hpc as _hpc
from . import lowering as _lowering

alpha_1sigma = 2 * scipy.stats.norm.cdf(-1.0)
"""
The alpha for the 1 sigma confidence level
"""


def percolation_graph(graph, spanning_cluster=True):
    """
    Prepare the (internal) percolation graph from a given graph

    Drop-in for percolate/percolate.py:31-100 (restated for networkx >= 2):
    strips the auxiliary nodes (node attribute ``'span'``) and returns the
    dict ``graph, spanning_cluster, auxiliary_node_attributes,
    spanning_sides, auxiliary_edge_attributes, perc_graph, num_nodes,
    num_edges`` that is splatted into the ``hpc.bond_*`` functions.

    Raises
    ------
    ValueError
        no auxiliary nodes, or not exactly two spanning sides
    """
    real = graph
    out = {'graph': graph, 'spanning_cluster': bool(spanning_cluster)}
    if spanning_cluster:
        aux = nx.get_node_attributes(graph, 'span')        # auxiliary node -> side
        if len(aux) == 0:
            raise ValueError(
                'Spanning cluster is to be detected, but no auxiliary nodes '
                'given.'
            )
        sides = list(set(aux.values()))
        if len(sides) != 2:
            raise ValueError(
                'Spanning cluster is to be detected, but auxiliary nodes '
                'of less or more than 2 types (sides) given.'
            )
        out.update(
            auxiliary_node_attributes=aux,
            spanning_sides=sides,
            auxiliary_edge_attributes=nx.get_edge_attributes(graph, 'span'),
        )
        # the real lattice: everything that is not an auxiliary node
        real = graph.subgraph([v for v in graph.nodes if v not in aux])
    out.update(perc_graph=real, num_nodes=real.number_of_nodes(),
               num_edges=real.number_of_edges())
    return out


# lowered forms of user graphs, keyed weakly by the graph object and validated by a fingerprint
# of its nodes, bonds and 'span' attributes (see lowering._LOWER_CACHE); the user's graph is
# never written to
import weakref

_PREPARED = weakref.WeakKeyDictionary()


def _prepare(graph, spanning_cluster):
    """LoweredGraph for ``graph`` (a networkx graph with auxiliary nodes, or a LoweredGraph)."""
    if isinstance(graph, _lowering.LoweredGraph):
        return _lowering.lower(graph, spanning_cluster=spanning_cluster)
    try:
        key = hash((bool(spanning_cluster), tuple(graph.nodes()), tuple(graph.edges()),
                    tuple(nx.get_node_attributes(graph, 'span').items()),
                    tuple(nx.get_edge_attributes(graph, 'span').items())))
        cached = _PREPARED.get(graph)
    except TypeError:
        key, cached = None, None
    if cached is not None and cached[0] == key:
        return cached[1]
    pg = percolation_graph(graph, spanning_cluster=spanning_cluster)
    lowered = _lowering.lower(
        pg['perc_graph'], spanning_cluster=spanning_cluster,
        auxiliary_node_attributes=pg.get('auxiliary_node_attributes'),
        auxiliary_edge_attributes=pg.get('auxiliary_edge_attributes'),
        spanning_sides=pg.get('spanning_sides'),
    )
    if key is not None:
        try:
            _PREPARED[graph] = (key, lowered)
        except TypeError:
            pass
    return lowered


def _context(lowered):
    return _native.context_for(lowered, _hpc._default_device())


def sample_states(
    graph, spanning_cluster=True, model='bond', copy_result=True
):
    '''
    Generate successive sample states of the percolation model

    Drop-in for percolate/percolate.py:103-356.  This is a generator function
    that adds one bond at a time and yields the cluster statistics as a
    dictionary.  Like the reference it yields the ``n == 0`` state BEFORE the
    bond order is drawn, and draws the order from NumPy's GLOBAL stream
    (``np.random.permutation``) when advanced past it; the run itself is then
    computed on the GPU in one batch.

    Yields
    ------
    ret : dict
        ``n``, ``N``, ``M``, ``edge`` (pair of nodes, only for ``n >= 1``),
        ``has_spanning_cluster`` (only if ``spanning_cluster``),
        ``max_cluster_size`` (int), ``moments`` (float64[5], moments k = 0..4
        of the cluster-size distribution without one largest cluster).
        With ``copy_result=False`` the same dict object is updated in place.

    Raises
    ------
    ValueError
        ``model != 'bond'``; no auxiliary nodes; not exactly two sides
    '''
    if model != 'bond':
        raise ValueError('Only bond percolation supported.')

    lowered = _prepare(graph, spanning_cluster)
    num_nodes = lowered.num_nodes
    num_edges = lowered.num_edges

    ret = dict()
    ret['n'] = 0
    ret['N'] = num_nodes
    ret['M'] = num_edges
    ret['max_cluster_size'] = 1
    ret['moments'] = np.ones(5) * (num_nodes - 1)
    if spanning_cluster:
        ret['has_spanning_cluster'] = False

    if copy_result:
        yield copy.deepcopy(ret)
    else:
        yield ret

    # permute edges: the reference's draw from the global stream
    # (percolate/percolate.py:270)
    perm_edges = np.random.permutation(num_edges)

    rows = _context(lowered).run_rows(
        1, _native.PERM_HOST, perm_edges.astype(np.int32)[None, :])[0]
    max_cluster_size = rows['max_cluster_size'].tolist()
    moments = rows['moments'].astype(np.float64)
    if spanning_cluster:
        has_spanning_cluster = rows['has_spanning_cluster'].tolist()

    for n in range(num_edges):
        ret['n'] = n + 1
        e = int(perm_edges[n])
        ret['edge'] = (lowered.label(int(lowered.eu[e])),
                       lowered.label(int(lowered.ev[e])))
        if spanning_cluster:
            ret['has_spanning_cluster'] = has_spanning_cluster[n + 1]
        ret['max_cluster_size'] = max_cluster_size[n + 1]
        ret['moments'][:] = moments[n + 1]

        if copy_result:
            yield copy.deepcopy(ret)
        else:
            yield ret


def single_run_arrays(spanning_cluster=True, **kwargs):
    r'''
    Generate statistics for a single run

    Drop-in for percolate/percolate.py:359-447: a stand-alone helper that
    evolves ``sample_states`` over all bonds and returns NumPy arrays
    ``max_cluster_size`` (float64[M+1]), ``moments`` (float64[5, M+1]),
    ``has_spanning_cluster`` (bool[M+1], only if ``spanning_cluster``),
    ``N`` and ``M``.
    '''
    states = sample_states(spanning_cluster=spanning_cluster,
                           **dict(kwargs, copy_result=False))
    first = next(states)
    num_nodes, num_edges = first['N'], first['M']
    largest = np.empty(num_edges + 1)
    moments = np.empty((5, num_edges + 1))
    spans = np.empty(num_edges + 1, dtype=bool) if spanning_cluster else None

    def store(state):
        assert (state['N'], state['M']) == (num_nodes, num_edges)
        n = state['n']
        largest[n] = state['max_cluster_size']
        moments[:, n] = state['moments']
        if spans is not None:
            spans[n] = state['has_spanning_cluster']

    store(first)
    for state in states:
        store(state)

    out = {'N': num_nodes, 'M': num_edges, 'max_cluster_size': largest,
           'moments': moments}
    if spans is not None:
        out['has_spanning_cluster'] = spans
    return out


def _microcanonical_average_spanning_cluster(has_spanning_cluster, alpha):
    r'''
    Compute the average number of runs that have a spanning cluster

    Drop-in for percolate/percolate.py:450-572: Bayesian posterior mean
    ``(k + 1) / (runs + 2)`` and the ``1 - alpha`` credible interval
    ``beta.ppf([alpha/2, 1 - alpha/2], k + 1, runs - k + 1)``.
    '''
    runs = has_spanning_cluster.size
    k = has_spanning_cluster.sum(dtype=float)
    tails = [alpha / 2, 1 - alpha / 2]
    return {
        'spanning_cluster': (k + 1) / (runs + 2),
        'spanning_cluster_ci': scipy.stats.beta.ppf(tails, k + 1, runs - k + 1),
    }


def _t_interval_or_point(mean, std, runs, alpha):
    """Student-t ``1 - alpha`` interval of a sample mean; the degenerate
    ``(mean, mean)`` when the sample standard deviation is zero
    (percolate/percolate.py:621-633, 691-703)."""
    if not std:
        return mean * np.ones(2)
    return scipy.stats.t.interval(1 - alpha, df=runs - 1, loc=mean,
                                  scale=std / np.sqrt(runs))


def _microcanonical_average_max_cluster_size(max_cluster_size, alpha):
    """
    Compute the average size of the largest cluster

    Drop-in for percolate/percolate.py:575-635: sample mean and Student-t
    ``1 - alpha`` interval; ``(mean, mean)`` when the sample std is zero.
    """
    mean = max_cluster_size.mean()
    std = max_cluster_size.std(ddof=1)
    return {
        'max_cluster_size': mean,
        'max_cluster_size_ci': _t_interval_or_point(
            mean, std, max_cluster_size.size, alpha),
    }


def _microcanonical_average_moments(moments, alpha):
    """
    Compute the average moments of the cluster size distributions

    Drop-in for percolate/percolate.py:638-705 (``moments`` has shape
    ``(runs, 5)``).
    """
    runs = moments.shape[0]
    mean = moments.mean(axis=0)
    std = moments.std(axis=0, ddof=1)
    ci = np.empty((5, 2))
    for k in range(5):
        ci[k] = _t_interval_or_point(mean[k], std[k], runs, alpha)
    return {'moments': mean, 'moments_ci': ci}


# credible intervals already evaluated, per (runs, alpha): [known[k], lo[k], hi[k]] for k = 0..runs
_BETA_TABLES = {}
_BETA_TABLES_MAX = 4


def _beta_interval_rows(k, runs, alpha):
    """``beta.ppf([alpha/2, 1 - alpha/2], k + 1, runs - k + 1)`` for every entry of the integer-valued
    array ``k`` (percolate/percolate.py:561-570), shape ``k.shape + (2,)``.

    The interval is a pure function of ``(k, runs, alpha)``: it is evaluated once per distinct k and
    kept in a table per ``(runs, alpha)``, so that repeated studies with the same number of runs
    (where the same counts come up again and again) do not pay scipy's root finder twice."""
    runs = int(runs)
    key = (runs, float(alpha))
    tab = _BETA_TABLES.get(key)
    if tab is None:
        if len(_BETA_TABLES) >= _BETA_TABLES_MAX:
            _BETA_TABLES.pop(next(iter(_BETA_TABLES)))
        tab = [np.zeros(runs + 1, dtype=bool), np.empty(runs + 1), np.empty(runs + 1)]
        _BETA_TABLES[key] = tab
    known, lo, hi = tab
    ki = np.asarray(k).astype(np.int64)
    need = np.unique(ki[~known[ki]])
    if need.size:
        q = scipy.stats.beta.ppf([[alpha / 2], [1 - alpha / 2]], need + 1.0, runs - need + 1.0)
        lo[need], hi[need] = q[0], q[1]
        known[need] = True
    return np.stack([lo[ki], hi[ki]], axis=-1)


def _interval(mean, var, runs, alpha):
    """Student-t interval per n from device means / exact variances
    (percolate/percolate.py:613-635, 681-705 vectorised over n).

    ``scipy.stats.t.interval(1 - alpha, df, loc, scale)`` evaluates
    ``t.ppf([alpha/2, 1 - alpha/2], df) * scale + loc``; the two quantiles
    depend on ``(alpha, runs)`` only, so they are taken once and the affine
    map is applied to all n -- the same floating-point operations the
    reference performs for every n separately.
    """
    std = np.sqrt(var)
    ci = np.empty(mean.shape + (2,))
    zero = (std == 0)
    with np.errstate(invalid='ignore'):
        q_lo, q_hi = scipy.stats.t.interval(1 - alpha, df=runs - 1)
        scale = std / np.sqrt(runs)
        ci[..., 0] = q_lo * scale + mean
        ci[..., 1] = q_hi * scale + mean
    ci[zero, 0] = mean[zero]
    ci[zero, 1] = mean[zero]
    return ci


class _MicrocanonicalAverages(object):
    """Iterator behind ``microcanonical_averages``.

    Follows the laziness of the reference's generator
    (percolate/percolate.py:834-896): argument validation on the first
    ``next()``; the ``n == 0`` ensemble is yielded before any random number
    is drawn; the ``runs`` bond orders are drawn from NumPy's global stream
    in run order when the iterator is advanced past ``n == 0``.  All runs are
    then swept and reduced on the GPU in one fused batch.
    """

    def __init__(self, graph, runs, spanning_cluster, model, alpha,
                 copy_result):
        self._args = (graph, runs, spanning_cluster, model, alpha, copy_result)
        self._state = 0          # 0 = nothing done, 1 = n == 0 yielded, 2 = arrays ready
        self._n = 0
        self._ret = dict()
        self._arrays = None

    def __iter__(self):
        return self

    def _validate(self):
        graph, runs, spanning_cluster, model, alpha, copy_result = self._args
        try:
            runs = int(runs)
        except Exception:
            raise ValueError("runs needs to be a positive integer")
        if runs <= 0:
            raise ValueError("runs needs to be a positive integer")
        try:
            alpha = float(alpha)
        except Exception:
            raise ValueError("alpha needs to be a float in the interval (0, 1)")
        if alpha <= 0.0 or alpha >= 1.0:
            raise ValueError("alpha needs to be a float in the interval (0, 1)")
        if model != 'bond':
            raise ValueError('Only bond percolation supported.')
        self._runs, self._alpha = runs, alpha
        self._spanning = bool(spanning_cluster)
        self._copy = copy_result
        self._lowered = _prepare(graph, spanning_cluster)

    def _compute(self):
        """Draw the bond orders like the reference and run the fused batch."""
        g = self._lowered
        runs, alpha = self._runs, self._alpha
        perms = np.empty((runs, g.num_edges), dtype=np.int32)
        for r in range(runs):
            perms[r] = np.random.permutation(g.num_edges)
        ctx = _context(g)
        ctx.reset_accumulators()
        ctx.run_fused(runs, _native.PERM_HOST, perms, _native.FUSE_MICRO)
        self._arrays = _arrays_on_device(ctx, runs, alpha, g.num_nodes,
                                         g.num_edges, self._spanning)
        ctx.reset_accumulators()

    def raw_arrays(self):
        """Un-normalised per-n arrays (used by microcanonical_averages_arrays
        to skip the per-n dictionaries when nothing has been consumed)."""
        if self._state == 0:
            self._validate()
            self._state = 1
        if self._arrays is None:
            self._compute()
        self._state = 3
        return self._arrays

    def __next__(self):
        if self._state == 0:
            self._validate()
            self._state = 1
            g = self._lowered
            ret = self._ret
            ret['n'] = 0
            ret['N'] = g.num_nodes
            ret['M'] = g.num_edges
            ret['max_cluster_size'] = 1.0
            ret['max_cluster_size_ci'] = np.ones(2)
            ret['moments'] = np.ones(5) * (g.num_nodes - 1)
            ret['moments_ci'] = np.ones((5, 2)) * (g.num_nodes - 1)
            if self._spanning:
                ret.update(_microcanonical_average_spanning_cluster(
                    np.zeros(self._runs), self._alpha))
            self._n = 1
            return copy.deepcopy(ret) if self._copy else ret
        if self._state == 3 or self._n > self._lowered.num_edges:
            if self._arrays is None and self._state != 3:
                # the reference's generators each draw their (empty) bond
                # order before stopping; nothing to compute for M == 0
                pass
            raise StopIteration
        if self._arrays is None:
            self._compute()
            self._state = 2
        a, n, ret = self._arrays, self._n, self._ret
        ret['n'] = n
        ret['max_cluster_size'] = a['max_cluster_size'][n]
        ret['max_cluster_size_ci'] = a['max_cluster_size_ci'][n].copy()
        ret['moments'] = a['moments'][:, n].copy()
        ret['moments_ci'] = a['moments_ci'][:, n].copy()
        if self._spanning:
            ret['spanning_cluster'] = a['spanning_cluster'][n]
            ret['spanning_cluster_ci'] = a['spanning_cluster_ci'][n].copy()
        self._n += 1
        return copy.deepcopy(ret) if self._copy else ret

    next = __next__


def _arrays_on_device(ctx, runs, alpha, num_nodes, num_edges, spanning, norm=1.0):
    """Per-n statistics of the runs folded into ``ctx``, each divided by ``norm``: the
    formulas of ``_arrays_from_device`` evaluated by ``pz_micro_arrays`` on the GPU
    (operation for operation, hence bit-identical); only the two Student-t quantiles,
    which depend on ``(alpha, runs)`` alone, and the beta table look-up stay on the host."""
    with np.errstate(invalid='ignore'):
        t_lo, t_hi = scipy.stats.t.interval(1 - alpha, df=runs - 1)
    k, largest, largest_ci, moments, moments_ci = ctx.micro_arrays(t_lo, t_hi, norm=norm)
    ret = {
        'max_cluster_size': largest, 'max_cluster_size_ci': largest_ci,
        'moments': moments, 'moments_ci': moments_ci,
    }
    if spanning:
        ret['spanning_cluster'] = (k + 1) / (runs + 2)
        ret['spanning_cluster_ci'] = _beta_interval_rows(k, runs, alpha)
    ret['M'] = num_edges
    ret['N'] = num_nodes
    return ret


def _arrays_from_device(mean, var, runs, alpha, num_nodes, num_edges, spanning):
    """Per-n statistics (NOT yet divided by N) from the device means / variances: the host
    form of ``_arrays_on_device`` (kept as its specification; the parity tests compare the two)."""
    ret = dict()
    ret['max_cluster_size'] = mean[1]
    ret['max_cluster_size_ci'] = _interval(mean[1], var[0], runs, alpha)
    ret['moments'] = mean[2:7]
    ret['moments_ci'] = _interval(mean[2:7], var[1:6], runs, alpha)
    if spanning:
        k = mean[0]
        # (k + 1) / (runs + 2) and the beta credible interval, evaluated once
        # per distinct k (percolate/percolate.py:561-570)
        ret['spanning_cluster'] = (k + 1) / (runs + 2)
        ret['spanning_cluster_ci'] = _beta_interval_rows(k, runs, alpha)
    ret['M'] = num_edges
    ret['N'] = num_nodes
    return ret


def microcanonical_averages(
    graph, runs=40, spanning_cluster=True, model='bond', alpha=alpha_1sigma,
    copy_result=True
):
    r'''
    Generate successive microcanonical percolation ensemble averages

    Drop-in for percolate/percolate.py:708-896.  Returns an iterator that
    yields, for every occupation number ``n = 0..M``, a dictionary with
    ``n, N, M``, ``max_cluster_size`` and ``max_cluster_size_ci`` (mean and
    Student-t interval over the runs), ``moments`` (float64[5]) and
    ``moments_ci`` (float64[5, 2]), and -- if ``spanning_cluster`` --
    ``spanning_cluster`` (Bayesian estimate ``(k+1)/(runs+2)``) and
    ``spanning_cluster_ci`` (beta credible interval).

    Raises
    ------
    ValueError
        ``runs`` is not a positive integer, ``alpha`` is not a float in (0, 1),
        or the errors of ``sample_states`` -- on the first ``next()``.
    '''
    return _MicrocanonicalAverages(graph, runs, spanning_cluster, model,
                                   alpha, copy_result)


def spanning_1d_chain(length):
    """
    Generate a linear chain with auxiliary nodes for spanning cluster detection

    Drop-in for percolate/percolate.py:899-928 (networkx >= 2 spelling).
    """
    length = int(length)
    chain = nx.path_graph(length + 2)           # nodes 0 .. length + 1
    ends = {0: (0, 1), 1: (length + 1, length)}   # side -> (auxiliary node, its neighbour)
    nx.set_node_attributes(chain, {aux: side for side, (aux, _) in ends.items()}, 'span')
    nx.set_edge_attributes(chain, {pair: side for side, pair in ends.items()}, 'span')
    return chain


def spanning_2d_grid(length):
    """
    Generate a square lattice with auxiliary nodes for spanning detection

    Drop-in for percolate/percolate.py:931-965 (networkx >= 2 spelling).
    """
    grid = nx.grid_2d_graph(length + 2, length)
    # columns 0 and length + 1 are auxiliary; each is tied to the lattice column next to it
    columns = {0: (0, 1), 1: (length + 1, length)}          # side -> (aux column, lattice column)
    nx.set_node_attributes(
        grid, {(aux, i): side for side, (aux, _) in columns.items() for i in range(length)}, 'span')
    nx.set_edge_attributes(
        grid, {((aux, i), (real, i)): side
               for side, (aux, real) in columns.items() for i in range(length)}, 'span')
    return grid


def microcanonical_averages_arrays(microcanonical_averages):
    """
    Compile microcanonical averages over all iteration steps into single arrays

    Drop-in for percolate/percolate.py:968-1064: stacks the per-n dictionaries
    and divides every key that does not contain ``'spanning_cluster'`` by the
    number of sites.
    """
    ret = dict()

    if isinstance(microcanonical_averages, _MicrocanonicalAverages) and \
            microcanonical_averages._state == 0:
        raw = microcanonical_averages.raw_arrays()
        num_edges, num_sites = raw['M'], raw['N']
        for key, value in raw.items():
            if len(key) <= 1:
                continue
            ret[key] = np.array(value, dtype=np.float64, copy=True)
        # the reference's layout: moments (5, M+1), moments_ci (5, M+1, 2)
    else:
        # any iterable of per-n dictionaries: collect, then stack
        cols = {}
        count = 0
        for n, avg in enumerate(microcanonical_averages):
            assert n == avg['n']
            if n == 0:
                num_edges, num_sites = avg['M'], avg['N']
                cols = {key: [] for key in avg if key.split('_ci')[0] in
                        ('max_cluster_size', 'spanning_cluster', 'moments')}
            for key, seq in cols.items():
                seq.append(np.array(avg[key], dtype=np.float64))
            count = n + 1
        assert count == num_edges + 1
        for key, seq in cols.items():
            stacked = np.stack(seq)                         # n first
            # the reference's layout: moments (5, M+1), moments_ci (5, M+1, 2)
            ret[key] = np.ascontiguousarray(np.moveaxis(stacked, 0, 1)) \
                if key.startswith('moments') else stacked

    # everything but the spanning probability is reported per site
    # (percolate/percolate.py:1056-1060)
    for key, value in ret.items():
        if 'spanning_cluster' not in key:
            value /= num_sites

    ret['M'] = num_edges
    ret['N'] = num_sites
    return ret


def _binomial_pmf(n, p):
    """
    Compute the binomial PMF according to Newman and Ziff

    Drop-in for percolate/percolate.py:1067-1109: mode-outward ratio
    recurrence, normalised by its sum.  Evaluated on the GPU (one thread per
    direction) in the reference's operation order.
    """
    n = int(n)
    ctx = _hpc._utility_context()
    pmf = ctx.set_ps(np.array([p], dtype=np.float64), want_pmf=True, M=n)
    return pmf[0]


def canonical_averages(ps, microcanonical_averages_arrays):
    """
    Compute the canonical cluster statistics from microcanonical statistics

    Drop-in for percolate/percolate.py:1112-1223: for every ``p`` the binomial
    weights are applied to every array key (means AND confidence bounds).  The
    weights and the ``[num_p x (M+1)] . [(M+1) x columns]`` contraction are
    computed on the GPU.
    """
    arrays = microcanonical_averages_arrays
    num_edges = arrays['M']
    num_p = ps.size
    ret = {'ps': ps, 'N': arrays['N'], 'M': num_edges}
    # output shapes: the n axis of every input array becomes the p axis
    for key, shape in (('max_cluster_size', (num_p,)), ('max_cluster_size_ci', (num_p, 2)),
                       ('spanning_cluster', (num_p,)), ('spanning_cluster_ci', (num_p, 2)),
                       ('moments', (5, num_p)), ('moments_ci', (5, num_p, 2))):
        if key in arrays:
            ret[key] = np.empty(shape)

    # gather every column of length M+1, contract once, scatter back
    columns = []
    targets = []
    for key, value in microcanonical_averages_arrays.items():
        if len(key) <= 1:
            continue
        if key in ['max_cluster_size', 'spanning_cluster']:
            columns.append(value)
            targets.append((key, ()))
        elif key in ['max_cluster_size_ci', 'spanning_cluster_ci']:
            for b in range(2):
                columns.append(value[:, b])
                targets.append((key, (slice(None), b)))
        elif key == 'moments':
            for k in range(5):
                columns.append(value[k])
                targets.append((key, (k,)))
        elif key == 'moments_ci':
            for k in range(5):
                for b in range(2):
                    columns.append(value[k, :, b])
                    targets.append((key, (k, slice(None), b)))
        else:
            raise NotImplementedError(
                '{}-dimensional array'.format(value.ndim)
            )

    ctx = _hpc._utility_context()
    ctx.set_ps(np.asarray(ps, dtype=np.float64), M=num_edges)
    out = ctx.convolve(np.ascontiguousarray(np.stack(columns), dtype=np.float64))
    for row, (key, index) in zip(out, targets):
        if index == ():
            ret[key][:] = row
        else:
            ret[key][index] = row

    return ret


def statistics(
    graph, ps, spanning_cluster=True, model='bond', alpha=alpha_1sigma, runs=40
):
    """
    Helper function to compute percolation statistics

    Drop-in for percolate/percolate.py:1226-1252:
    ``microcanonical_averages`` -> ``microcanonical_averages_arrays`` ->
    ``canonical_averages``.
    """
    per_n = microcanonical_averages_arrays(microcanonical_averages(
        graph=graph, runs=runs, spanning_cluster=spanning_cluster, model=model, alpha=alpha))
    return canonical_averages(ps, per_n)
