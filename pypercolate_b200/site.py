"""
Site percolation on the bond-percolation sweep.

The reference lists site percolation as a TODO (percolate/__init__.py:68-72,
issue #5) and has no behaviour to match; the semantics here are those of the
Newman-Ziff paper the reference implements (Phys. Rev. E 64, 016706): the
sites of the graph are occupied one at a time in a random order, a cluster is
a connected component of the subgraph induced by the occupied sites, and after
every site the run reports what the bond runs report -- the largest cluster,
the spanning flag and the k = 0..4 moments of the cluster sizes without one
largest cluster.

No new kernel: occupying a site activates the bonds to its occupied
neighbours, so a site run IS a bond run with a derived bond order -- the bonds
sorted by the step at which their later endpoint is occupied (bonds activated
by the same site in any order: the state after the site is the same).  The
bond sweep treats every node as a cluster of size 1 from the start; the
``u = N - n`` unoccupied sites are exactly the extra singletons, each adding 1
to every moment, so after n >= 1 sites

    max_cluster_size = bond max           moments[k] = bond moments[k] - u

(mod 2^64, the arithmetic of the reference's uint64 moments) read at the row
of the bond run that follows the last bond of site n.  The derived orders are
built on the host (numpy), the sweep runs on the GPU through the same C-ABI
call as ``bond_microcanonical_statistics`` (caller-supplied bond orders,
validated on the device).
"""

import numpy as np

from . import _native, hpc

_SITE_FIELDS = (
    ('n', 'uint32', False),
    ('site', 'uint32', False),
    ('has_spanning_cluster', 'bool', True),
    ('max_cluster_size', 'uint32', False),
    ('moments', '(5,)uint64', False),
)


def site_microcanonical_statistics_dtype(spanning_cluster=True):
    """Structured dtype of a site run: the bond rows' fields with ``site`` (index of the site
    occupied last, in the order of ``perc_graph``'s percolating nodes) in the place of ``edge``."""
    return hpc._ndarray_dtype(_SITE_FIELDS, spanning_cluster)


def derived_bond_order(num_nodes, eu, ev, site_order):
    """Bond order of the bond run that is equivalent to occupying the sites in ``site_order``.

    Returns ``(bond_order, bonds_after)``: ``bonds_after[n]`` = number of bonds present once
    the first ``n`` sites are occupied (``bonds_after[0] == 0``)."""
    site_order = np.asarray(site_order, dtype=np.int64)
    if site_order.shape != (num_nodes,) or not np.array_equal(np.sort(site_order), np.arange(num_nodes)):
        raise ValueError("site_order must be a permutation of the sites")
    rank = np.empty(num_nodes, dtype=np.int64)
    rank[site_order] = np.arange(num_nodes)
    act = np.maximum(rank[np.asarray(eu, dtype=np.int64)], rank[np.asarray(ev, dtype=np.int64)])
    order = np.argsort(act, kind='stable').astype(np.int32)
    bonds_after = np.zeros(num_nodes + 1, dtype=np.int64)
    bonds_after[1:] = np.searchsorted(act[order], np.arange(num_nodes), side='right')
    return order, bonds_after


def bond_graph_for_sites(lowered):
    """The graph the bond sweep runs on.  A site that touches BOTH spanning sides joins them by
    itself; in a bond run that means "spanning from the first merge on" (the reference's flag is
    re-evaluated on merges only), whether the site is occupied or not.  Such sites take part in
    the bond run without their side bits; ``site_rows_from_bond_rows`` lets them span from the
    moment they are occupied."""
    from . import lowering
    if lowered.side_mask is None or not np.any(lowered.side_mask == 3):
        return lowered
    cached = lowered.__dict__.get('_site_graph')
    if cached is None:
        mask = lowered.side_mask.copy()
        mask[mask == 3] = 0
        cached = lowering.LoweredGraph(lowered.num_nodes, lowered.eu, lowered.ev, side_mask=mask,
                                       preconnected=lowered.preconnected, node_labels=lowered.node_labels)
        lowered._site_graph = cached        # (one device context per graph, not per call)
    return cached


def site_rows_from_bond_rows(bond_rows, site_order, bonds_after, side_mask, spanning_cluster=True):
    """Rows of a site run from the rows of the equivalent bond run (see the module docstring)."""
    N = len(site_order)
    out = np.zeros(N + 1, dtype=site_microcanonical_statistics_dtype(spanning_cluster))
    picked = bond_rows[bonds_after]
    n = np.arange(N + 1, dtype=np.uint64)
    out['n'] = n
    out['site'][1:] = site_order
    out['max_cluster_size'][1:] = picked['max_cluster_size'][1:]
    with np.errstate(over='ignore'):
        out['moments'][1:] = picked['moments'][1:] - (np.uint64(N) - n[1:])[:, None]
    if spanning_cluster:
        span = picked['has_spanning_cluster'].copy()
        span[0] = False
        if side_mask is not None:
            # a site that touches both sides spans from the moment it is occupied (the bond run
            # was made without its side bits, see bond_graph_for_sites)
            both = np.asarray(side_mask)[np.asarray(site_order)] == 3
            if both.any():
                span[1 + int(np.argmax(both)):] = True
        out['has_spanning_cluster'] = span
    return out


def site_microcanonical_statistics(
    perc_graph, num_nodes, num_edges, seed, spanning_cluster=True,
    auxiliary_node_attributes=None, auxiliary_edge_attributes=None,
    spanning_sides=None, **kwargs
):
    """
    Evolve a single SITE-percolation run over all ``num_nodes + 1`` site occupation numbers

    Same arguments as ``hpc.bond_microcanonical_statistics``.  The site order is
    ``numpy.random.RandomState(seed).permutation(num_nodes)`` over the percolating nodes in the
    graph's node order.
    """
    return site_microcanonical_statistics_batch(
        perc_graph, num_nodes, num_edges, [seed], spanning_cluster, auxiliary_node_attributes,
        auxiliary_edge_attributes, spanning_sides, **kwargs)[0]


def site_microcanonical_statistics_batch(
    perc_graph, num_nodes, num_edges, seeds, spanning_cluster=True,
    auxiliary_node_attributes=None, auxiliary_edge_attributes=None,
    spanning_sides=None, **kwargs
):
    """``site_microcanonical_statistics`` for many seeds in one device batch: array of shape
    ``(len(seeds), num_nodes + 1)``."""
    lowered = hpc._lower(perc_graph, spanning_cluster, auxiliary_node_attributes,
                         auxiliary_edge_attributes, spanning_sides)
    if lowered.num_nodes != num_nodes or lowered.num_edges != num_edges:
        raise ValueError('num_nodes / num_edges do not match perc_graph')
    if lowered.preconnected:
        raise ValueError('site percolation needs a graph whose sites start unconnected')
    seeds = list(seeds)
    N, M = lowered.num_nodes, lowered.num_edges
    site_orders = np.empty((len(seeds), N), dtype=np.int64)
    perms = np.empty((len(seeds), M), dtype=np.int32)
    after = np.empty((len(seeds), N + 1), dtype=np.int64)
    for r, s in enumerate(seeds):
        site_orders[r] = np.random.RandomState(seed=s).permutation(N)
        perms[r], after[r] = derived_bond_order(N, lowered.eu, lowered.ev, site_orders[r])
    device = kwargs.get('device')
    ctx = _native.context_for(bond_graph_for_sites(lowered),
                              hpc._default_device() if device is None else device)
    rows = ctx.run_rows(len(seeds), _native.PERM_HOST, perms)
    out = np.empty((len(seeds), N + 1), dtype=site_microcanonical_statistics_dtype(spanning_cluster))
    for r in range(len(seeds)):
        out[r] = site_rows_from_bond_rows(rows[r], site_orders[r], after[r], lowered.side_mask,
                                          spanning_cluster)
    return out
