"""Lowering of networkx percolation graphs to flat arrays (host logic, no GPU)."""
import networkx as nx
import numpy as np
import pytest

from conftest import HPC_FIXTURES, load_golden
from pypercolate_b200 import lowering, percolate


def lower_graph(graph, spanning=True):
    pg = percolate.percolation_graph(graph, spanning_cluster=spanning)
    return lowering.lower(pg['perc_graph'], spanning, pg.get('auxiliary_node_attributes'),
                          pg.get('auxiliary_edge_attributes'), pg.get('spanning_sides')), pg


@pytest.mark.parametrize("L", [1, 2, 3, 5, 8, 13])
def test_closed_form_2d_grid_equals_networkx_lowering(L):
    low, pg = lower_graph(percolate.spanning_2d_grid(L))
    closed = lowering.lowered_spanning_2d_grid(L)
    assert low.num_nodes == closed.num_nodes == L * L
    assert low.num_edges == closed.num_edges == 2 * L * (L - 1)
    assert np.array_equal(low.eu, closed.eu) and np.array_equal(low.ev, closed.ev)
    assert np.array_equal(low.side_mask, closed.side_mask)
    assert closed.edges() == list(pg['perc_graph'].edges())
    assert list(closed.nodes_iter()) == list(pg['perc_graph'].nodes())


@pytest.mark.parametrize("L", [1, 2, 10])
def test_closed_form_chain_equals_networkx_lowering(L):
    low, pg = lower_graph(percolate.spanning_1d_chain(L))
    closed = lowering.lowered_spanning_1d_chain(L)
    assert np.array_equal(low.eu, closed.eu) and np.array_equal(low.ev, closed.ev)
    assert np.array_equal(low.side_mask, closed.side_mask)
    assert low.preconnected == closed.preconnected or L == 1


def test_chain_of_one_node_touches_both_sides():
    low, _ = lower_graph(percolate.spanning_1d_chain(1))
    assert low.num_nodes == 1 and low.num_edges == 0
    assert low.side_mask.tolist() == [3] and low.preconnected


@pytest.mark.parametrize("L", [2, 3, 4])
def test_closed_form_3d_grid_equals_networkx(L):
    g = nx.grid_graph([L, L, L])
    closed = lowering.lowered_spanning_3d_grid(L)
    nodes = list(g.nodes())
    assert [closed.label(i) for i in range(closed.num_nodes)] == nodes
    assert closed.edges() == list(g.edges())
    assert closed.num_edges == 3 * L * L * (L - 1)
    assert set(np.flatnonzero(closed.side_mask & 1)) == {i for i, n in enumerate(nodes) if n[0] == 0}
    assert set(np.flatnonzero(closed.side_mask & 2)) == {i for i, n in enumerate(nodes) if n[0] == L - 1}


@pytest.mark.parametrize("name", HPC_FIXTURES)
def test_fixture_graphs_are_consistent(name):
    d = load_golden(name)
    assert d['eu'].shape == d['ev'].shape == (int(d['M']),)
    if int(d['spanning']):
        assert d['side_mask'].shape == (int(d['N']),)


def test_percolation_graph_keys_and_errors():
    pg = percolate.percolation_graph(percolate.spanning_2d_grid(3))
    assert set(pg) == {'graph', 'spanning_cluster', 'auxiliary_node_attributes', 'spanning_sides',
                       'auxiliary_edge_attributes', 'perc_graph', 'num_nodes', 'num_edges'}
    assert pg['num_nodes'] == 9 and pg['num_edges'] == 12 and sorted(pg['spanning_sides']) == [0, 1]
    with pytest.raises(ValueError):
        percolate.percolation_graph(nx.Graph())
    one_sided = nx.path_graph(4)
    one_sided.nodes[0]['span'] = 0
    with pytest.raises(ValueError):
        percolate.percolation_graph(one_sided)
    pg2 = percolate.percolation_graph(nx.path_graph(4), spanning_cluster=False)
    assert pg2['num_nodes'] == 4 and 'spanning_sides' not in pg2


def test_lower_rejects_wrong_number_of_sides():
    g = nx.path_graph(3)
    with pytest.raises(ValueError):
        lowering.lower(g, True, {}, {}, [0])
    with pytest.raises(ValueError):
        lowering.lower(g, True, {}, {}, [0, 1, 2])


def test_lowered_graph_validates_endpoints():
    with pytest.raises(ValueError):
        lowering.LoweredGraph(3, [0, 1], [1, 3])
    with pytest.raises(ValueError):
        lowering.LoweredGraph(3, [0, 1], [1, 2], side_mask=[0, 1])


def test_preconnected_detection():
    g = nx.path_graph(6)
    g.add_node('a', span=0)
    g.add_node('b', span=1)
    g.add_edge('a', 2, span=0)
    g.add_edge('b', 2, span=1)
    low, _ = lower_graph(g)
    assert low.preconnected and low.side_mask[2] == 3
    # an auxiliary edge joining auxiliary nodes of both sides
    h = nx.path_graph(4)
    h.add_node('a', span=0)
    h.add_node('b', span=1)
    h.add_edge('a', 0, span=0)
    h.add_edge('b', 3, span=1)
    h.add_edge('a', 'b', span=0)
    low, _ = lower_graph(h)
    assert low.preconnected


# -- caches and copies (the user's graph object is never written to) -------------

def test_lowering_follows_in_place_mutation_of_the_graph():
    g = percolate.spanning_2d_grid(4)
    a = percolate._prepare(g, True)
    assert percolate._prepare(g, True) is a                       # unchanged graph: cached
    # rewire one bond, same node and bond counts
    g.remove_edge((1, 0), (1, 1))
    g.add_edge((1, 0), (3, 3))
    b = percolate._prepare(g, True)
    assert b is not a and b.num_edges == a.num_edges
    assert b.fingerprint() != a.fingerprint()
    assert ((1, 0), (3, 3)) in b.edges() and ((1, 0), (1, 1)) not in b.edges()
    # move an auxiliary bond to the other side: same counts again
    edge = next(e for e, s in nx.get_edge_attributes(g, 'span').items() if s == 0)
    g.edges[edge]['span'] = 1
    c = percolate._prepare(g, True)
    assert c is not b and not np.array_equal(c.side_mask, b.side_mask)


def test_lower_cache_keyed_by_content_not_identity():
    pg = percolate.percolation_graph(percolate.spanning_2d_grid(3))
    args = (pg['perc_graph'], True, pg['auxiliary_node_attributes'],
            pg['auxiliary_edge_attributes'], pg['spanning_sides'])
    a = lowering.lower(*args)
    assert lowering.lower(*args) is a
    # equal content in fresh dict objects still hits; changed content does not
    assert lowering.lower(args[0], True, dict(args[2]), dict(args[3]), list(args[4])) is a
    flipped = {e: 1 - s for e, s in args[3].items()}
    b = lowering.lower(args[0], True, args[2], flipped, args[4])
    # (the auxiliary nodes keep their side, their bonds now carry the other one: both hubs merge)
    assert b is not a and b.preconnected and not a.preconnected


def test_graphs_still_pickle_and_deepcopy_after_lowering():
    import copy
    import pickle
    g = percolate.spanning_2d_grid(3)
    low = percolate._prepare(g, True)
    low._handles[0] = object()              # stands in for a ctypes device context
    assert '_pz_percolation' not in g.__dict__ and '_pz_lowered' not in g.__dict__
    g2 = pickle.loads(pickle.dumps(g))
    assert sorted(g2.edges()) == sorted(g.edges())
    copy.deepcopy(g)
    pg = percolate.percolation_graph(g)
    lowering.lower(pg['perc_graph'], True, pg['auxiliary_node_attributes'],
                   pg['auxiliary_edge_attributes'], pg['spanning_sides'])
    copy.deepcopy(pg)
    for big in (low, lowering.lowered_spanning_2d_grid(70), lowering.lowered_spanning_3d_grid(5)):
        big._handles[0] = object()
        for clone in (pickle.loads(pickle.dumps(big)), copy.deepcopy(big)):
            assert clone._handles == {} and clone.fingerprint() == big.fingerprint()
            assert clone.label(clone.num_nodes - 1) == big.label(big.num_nodes - 1)
            assert clone.edges()[:5] == big.edges()[:5]
