"""
Lowering of a percolation graph to the flat arrays the CUDA kernels consume.

The reference keeps the substrate as a ``networkx`` graph and walks it bond by
bond (percolate/hpc.py:205,246,254-262).  Here the graph is lowered ONCE to

* ``eu[M]``, ``ev[M]``  int32 endpoints of bond ``e`` in ``perc_graph.edges()``
  order, node ids being positions in ``perc_graph.nodes()`` order -- so the
  ``edge`` field of a row (percolate/hpc.py:254-256) keeps its meaning;
* ``side_mask[N]``  uint8, bit ``s`` set iff the auxiliary structure of
  spanning side ``s`` touches real node ``x`` (percolate/hpc.py:225-243);
* ``preconnected``  the auxiliary structure alone already joins both sides
  (then the flag reads True from the first merge on, percolate/hpc.py:266-274).

Closed-form generators build the same arrays for the square / cubic / chain
lattices without networkx (the reference's own stated limit is the ~1 GiB per
10^6 nodes networkx graph, docs/pypercolate-hpc.rst:11-22).
"""

import numpy as np


class LoweredGraph(object):
    """Flat description of a percolation graph.

    Duck-types the two members of ``perc_graph`` the reference's sweep touches
    (``edges()`` and ``nodes_iter()``, percolate/hpc.py:205,246), so an instance
    can be passed wherever the reference expects ``perc_graph``.
    """

    def __init__(self, num_nodes, eu, ev, side_mask=None, preconnected=False,
                 node_labels=None):
        self.num_nodes = int(num_nodes)
        self.eu = np.ascontiguousarray(eu, dtype=np.int32)
        self.ev = np.ascontiguousarray(ev, dtype=np.int32)
        if self.eu.shape != self.ev.shape or self.eu.ndim != 1:
            raise ValueError("eu and ev must be 1-d arrays of equal length")
        self.num_edges = int(self.eu.size)
        if self.num_edges and (
                min(self.eu.min(), self.ev.min()) < 0 or
                max(self.eu.max(), self.ev.max()) >= self.num_nodes):
            raise ValueError("edge endpoint out of range")
        if side_mask is not None:
            side_mask = np.ascontiguousarray(side_mask, dtype=np.uint8)
            if side_mask.shape != (self.num_nodes,):
                raise ValueError("side_mask must have one entry per node")
        self.side_mask = side_mask
        self.preconnected = bool(preconnected)
        self.node_labels = node_labels      # list of original node objects or None
        self.lattice = None                 # ('grid2d' | 'grid3d', L): labels in closed form
        self._handles = {}                  # device id -> native handle (set by _native)

    # device contexts hold ctypes pointers: they never travel with a copy or a pickle
    def __getstate__(self):
        state = dict(self.__dict__)
        state['_handles'] = {}
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)
        self._handles = {}

    # -- networkx-like surface ------------------------------------------------
    def label(self, i):
        if self.node_labels is not None:
            return self.node_labels[i]
        if self.lattice is not None:
            kind, L = self.lattice
            if kind == 'grid2d':
                return (i // L + 1, i % L)
            return (i // (L * L), (i // L) % L, i % L)
        return i

    def edges(self):
        return [(self.label(int(u)), self.label(int(v)))
                for u, v in zip(self.eu, self.ev)]

    def nodes_iter(self):
        return iter(range(self.num_nodes) if self.node_labels is None
                    else self.node_labels)

    def nodes(self):
        return list(self.nodes_iter())

    def number_of_nodes(self):
        return self.num_nodes

    def number_of_edges(self):
        return self.num_edges

    def without_spanning(self):
        g = LoweredGraph(self.num_nodes, self.eu, self.ev, None, False,
                         self.node_labels)
        g.lattice = self.lattice
        return g

    def fingerprint(self):
        """Content hash of everything the kernels consume."""
        import hashlib
        h = hashlib.blake2b(digest_size=16)
        h.update(np.int64([self.num_nodes, self.num_edges, int(self.preconnected)]).tobytes())
        h.update(self.eu.tobytes())
        h.update(self.ev.tobytes())
        h.update(b'-' if self.side_mask is None else self.side_mask.tobytes())
        return h.hexdigest()


class _HostUnionFind(object):
    """Tiny dict union-find for the auxiliary structure only (a few nodes)."""

    def __init__(self):
        self.parent = {}

    def find(self, x):
        p = self.parent.setdefault(x, x)
        while p != self.parent[p]:
            self.parent[p] = self.parent[self.parent[p]]
            p = self.parent[p]
        self.parent[x] = p
        return p

    def union(self, *xs):
        r = self.find(xs[0])
        for x in xs[1:]:
            self.parent[self.find(x)] = r


# Lowered forms of graphs seen before: a side table keyed weakly by the graph object (the caller's
# object is never written to, so it still pickles and deep-copies), validated by a fingerprint of
# everything the lowering reads -- a graph mutated in place (a rewired bond, a changed 'span'
# attribute) is lowered again.
import weakref

_LOWER_CACHE = weakref.WeakKeyDictionary()


def _graph_fingerprint(perc_graph, spanning_cluster, auxiliary_node_attributes,
                       auxiliary_edge_attributes, spanning_sides):
    def attr_items(d):
        return None if d is None else tuple(d.items())
    return hash((bool(spanning_cluster), tuple(perc_graph.nodes()), tuple(perc_graph.edges()),
                 attr_items(auxiliary_node_attributes), attr_items(auxiliary_edge_attributes),
                 None if spanning_sides is None else tuple(spanning_sides)))


def lower(perc_graph, spanning_cluster=True, auxiliary_node_attributes=None,
          auxiliary_edge_attributes=None, spanning_sides=None):
    """Lower ``perc_graph`` (+ the auxiliary attributes of
    ``percolation_graph()``, percolate/percolate.py:55-79) to a ``LoweredGraph``.

    Raises the reference's ``ValueError`` when ``spanning_sides`` does not hold
    exactly two sides (percolate/hpc.py:197-202).
    """
    if isinstance(perc_graph, LoweredGraph):
        # already lowered: the two sides are the bits of its side mask
        if spanning_cluster and perc_graph.side_mask is None:
            raise ValueError(
                'Spanning cluster is to be detected, but auxiliary nodes '
                'of less or more than 2 types (sides) given.'
            )
        return perc_graph if spanning_cluster else (
            perc_graph if perc_graph.side_mask is None
            else perc_graph.without_spanning())

    if spanning_cluster:
        if spanning_sides is None or len(spanning_sides) != 2:
            raise ValueError(
                'Spanning cluster is to be detected, but auxiliary nodes '
                'of less or more than 2 types (sides) given.'
            )

    try:
        key = _graph_fingerprint(perc_graph, spanning_cluster, auxiliary_node_attributes,
                                 auxiliary_edge_attributes, spanning_sides)
        cached = _LOWER_CACHE.get(perc_graph)
    except TypeError:                      # unhashable node objects / not weak-referenceable
        key, cached = None, None
    if cached is not None and cached[0] == key:
        return cached[1]

    nodes = list(perc_graph.nodes())
    index = {node: i for i, node in enumerate(nodes)}
    edges = list(perc_graph.edges())
    M = len(edges)
    eu = np.fromiter((index[u] for u, _ in edges), dtype=np.int32, count=M)
    ev = np.fromiter((index[v] for _, v in edges), dtype=np.int32, count=M)

    side_mask = None
    preconnected = False
    if spanning_cluster:
        # percolate/hpc.py:228-243: every auxiliary node of a side is merged
        # into one hub; every auxiliary edge joins (hub of the EDGE's side,
        # both endpoints)
        sides = list(spanning_sides)
        uf = _HostUnionFind()
        hubs = [('hub', 0), ('hub', 1)]

        def place(node):
            if node in auxiliary_node_attributes:
                return hubs[sides.index(auxiliary_node_attributes[node])]
            return ('real', index[node])

        touched = set()
        for (edge, edge_side) in auxiliary_edge_attributes.items():
            hub = hubs[sides.index(edge_side)] if edge_side in sides else None
            if hub is None:
                raise KeyError(edge_side)
            members = [place(x) for x in edge]
            uf.union(hub, *members)
            touched.update(m for m in members if m[0] == 'real')
        roots = [uf.find(h) for h in hubs]
        preconnected = roots[0] == roots[1]
        side_mask = np.zeros(len(nodes), dtype=np.uint8)
        for m in touched:
            r = uf.find(m)
            side_mask[m[1]] = (1 if r == roots[0] else 0) | (2 if r == roots[1] else 0)

    lowered = LoweredGraph(len(nodes), eu, ev, side_mask, preconnected, nodes)
    if key is not None:
        try:
            _LOWER_CACHE[perc_graph] = (key, lowered)
        except TypeError:
            pass
    return lowered


# -- closed-form lattices -----------------------------------------------------

def lowered_spanning_2d_grid(length):
    """Arrays equal to ``lower(percolation_graph(spanning_2d_grid(L)))``
    (percolate/percolate.py:954-963) without building a networkx graph.

    Node (i, j), i = 1..L, j = 0..L-1 has id (i-1)*L + j; bonds are emitted per
    node in id order, first to (i+1, j) then to (i, j+1) -- the iteration
    order of ``grid_2d_graph(L+2, L)`` restricted to the real nodes.
    Side 0 touches row i = 1, side 1 touches row i = L.
    """
    L = int(length)
    ids = np.arange(L * L, dtype=np.int64).reshape(L, L)
    down = np.full((L, L), -1, dtype=np.int64)
    right = np.full((L, L), -1, dtype=np.int64)
    down[:-1, :] = ids[1:, :]
    right[:, :-1] = ids[:, 1:]
    tgt = np.stack([down, right], axis=2).reshape(-1)
    src = np.repeat(ids.reshape(-1), 2)
    keep = tgt >= 0
    side_mask = np.zeros(L * L, dtype=np.uint8)
    side_mask[ids[0, :]] |= 1
    side_mask[ids[-1, :]] |= 2
    labels = [(i + 1, j) for i in range(L) for j in range(L)] if L <= 64 else None
    g = LoweredGraph(L * L, src[keep], tgt[keep], side_mask, False, labels)
    if labels is None:
        g.lattice = ('grid2d', L)
    return g


def lowered_spanning_1d_chain(length):
    """``spanning_1d_chain`` (percolate/percolate.py:921-927): real nodes 1..L,
    side 0 touches node 1, side 1 touches node L."""
    L = int(length)
    eu = np.arange(0, L - 1, dtype=np.int32)
    side_mask = np.zeros(L, dtype=np.uint8)
    side_mask[0] |= 1
    side_mask[L - 1] |= 2
    return LoweredGraph(L, eu, eu + 1, side_mask, False, list(range(1, L + 1)))


def lowered_spanning_3d_grid(length):
    """Simple-cubic lattice L^3 as a general edge list: id (x*L + y)*L + z,
    bonds per node in id order to +x, +y, +z (the iteration order of
    ``networkx.grid_graph([L, L, L])``); side 0 = plane x = 0, side 1 = plane
    x = L-1 (BASELINE config 5)."""
    L = int(length)
    ids = np.arange(L ** 3, dtype=np.int64).reshape(L, L, L)
    nbr = np.full((L, L, L, 3), -1, dtype=np.int64)
    nbr[:-1, :, :, 0] = ids[1:, :, :]
    nbr[:, :-1, :, 1] = ids[:, 1:, :]
    nbr[:, :, :-1, 2] = ids[:, :, 1:]
    tgt = nbr.reshape(-1)
    src = np.repeat(ids.reshape(-1), 3)
    keep = tgt >= 0
    side_mask = np.zeros(L ** 3, dtype=np.uint8)
    side_mask[ids[0].reshape(-1)] |= 1
    side_mask[ids[-1].reshape(-1)] |= 2
    g = LoweredGraph(L ** 3, src[keep], tgt[keep], side_mask, False, None)
    g.lattice = ('grid3d', L)
    return g
