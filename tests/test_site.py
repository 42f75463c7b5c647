"""Site percolation (a TODO of the reference, percolate/__init__.py:68-72): the
derivation of a site run from the equivalent bond run, checked against a brute-force evaluation of
the definition (connected components of the subgraph induced by the occupied sites)."""
import numpy as np
import pytest

from pypercolate_b200 import lowering, site


def brute_force_site_rows(g, site_order, spanning):
    import networkx as nx
    N = g.num_nodes
    G = nx.Graph()
    adj = [[] for _ in range(N)]
    for u, v in zip(g.eu.tolist(), g.ev.tolist()):
        adj[u].append(v); adj[v].append(u)
    rows = []
    occupied = set()
    for n in range(N + 1):
        if n:
            s = int(site_order[n - 1])
            occupied.add(s)
            G.add_node(s)
            for t in adj[s]:
                if t in occupied:
                    G.add_edge(s, t)
        comps = [c for c in nx.connected_components(G)]
        sizes = sorted((len(c) for c in comps), reverse=True)
        mx = sizes[0] if sizes else 0
        rest = sizes[1:]
        mom = [sum(x ** k for x in rest) % 2 ** 64 for k in range(5)]
        span = False
        if spanning:
            for c in comps:
                m = 0
                for x in c:
                    m |= int(g.side_mask[x])
                if m == 3:
                    span = True
        rows.append((n, mx, mom, span))
    return rows


def check(g, seed, bond_engine):
    N, M = g.num_nodes, g.num_edges
    spanning = g.side_mask is not None
    order = np.random.RandomState(seed).permutation(N)
    perm, after = site.derived_bond_order(N, g.eu, g.ev, order)
    assert np.array_equal(np.sort(perm), np.arange(M)) and after[0] == 0 and after[-1] == M
    assert np.all(np.diff(after) >= 0)
    got = site.site_rows_from_bond_rows(bond_engine(g, perm), order, after, g.side_mask, spanning)
    want = brute_force_site_rows(g, order, spanning)
    assert got['n'].tolist() == [w[0] for w in want]
    assert got['site'][1:].tolist() == order.tolist()
    assert got['max_cluster_size'].tolist() == [w[1] for w in want]
    assert got['moments'].tolist() == [w[2] for w in want]
    if spanning:
        assert got['has_spanning_cluster'].tolist() == [w[3] for w in want]


def oracle_engine(g, perm):
    from oracle import oracle
    g = site.bond_graph_for_sites(g)
    return oracle.sweep_rows(g.num_nodes, g.num_edges, g.eu, g.ev, g.side_mask, g.preconnected, perm)


@pytest.mark.parametrize("seed", [0, 1, 7])
def test_site_rows_from_bond_rows_against_brute_force(seed):
    check(lowering.lowered_spanning_2d_grid(7), seed, oracle_engine)
    check(lowering.lowered_spanning_3d_grid(4), seed, oracle_engine)
    check(lowering.lowered_spanning_1d_chain(9), seed, oracle_engine)
    rng = np.random.RandomState(100 + seed)
    eu, ev = rng.randint(0, 30, 70), rng.randint(0, 30, 70)          # self loops and multi-edges included
    check(lowering.LoweredGraph(30, eu, ev), seed, oracle_engine)
    mask = np.zeros(30, dtype=np.uint8); mask[:4] = 1; mask[-4:] = 2; mask[10] = 3
    check(lowering.LoweredGraph(30, eu, ev, side_mask=mask), seed, oracle_engine)


def test_derived_bond_order_rejects_bad_site_orders():
    g = lowering.lowered_spanning_2d_grid(3)
    with pytest.raises(ValueError):
        site.derived_bond_order(g.num_nodes, g.eu, g.ev, np.zeros(g.num_nodes, dtype=int))


@pytest.mark.gpu
def test_site_percolation_on_the_gpu_against_brute_force():
    import percolate
    graph = percolate.spanning_2d_grid(12)
    pg = percolate.percolate.percolation_graph(graph)
    rows = percolate.site.site_microcanonical_statistics_batch(seeds=[3, 4, 5], **pg)
    g = lowering.lower(**{k: v for k, v in pg.items() if k not in ('num_nodes', 'num_edges', 'graph')})
    assert rows.shape == (3, pg['num_nodes'] + 1)
    for r, seed in enumerate([3, 4, 5]):
        order = np.random.RandomState(seed).permutation(g.num_nodes)
        want = brute_force_site_rows(g, order, True)
        assert rows[r]['max_cluster_size'].tolist() == [w[1] for w in want]
        assert rows[r]['moments'].tolist() == [w[2] for w in want]
        assert rows[r]['has_spanning_cluster'].tolist() == [w[3] for w in want]
    one = percolate.site.site_microcanonical_statistics(seed=4, **pg)
    assert np.array_equal(one, rows[1])
    # the site threshold of the square lattice is 0.5927: spanning sets in well above the bond value
    big = percolate.spanning_2d_grid(64)
    pgb = percolate.percolate.percolation_graph(big)
    many = percolate.site.site_microcanonical_statistics_batch(seeds=range(40), **pgb)
    first = np.array([np.argmax(m['has_spanning_cluster']) for m in many]) / float(pgb['num_nodes'])
    assert 0.55 < first.mean() < 0.64
