"""
TEST INFRASTRUCTURE: generate ``tests/golden/*.npz`` by running the UNMODIFIED
reference (``/root/reference/percolate/{percolate,hpc}.py`` through
``oracle/ref_shim.py``).  Runs only in the build container (the reference tree
is not on the GPU box); the fixtures it writes are committed.

    python oracle/make_golden.py                     # everything (minutes)
    python oracle/make_golden.py --only=big_grid256  # just the named fixtures

Every fixture stores its inputs (lowered graph arrays, seeds, ps) next to the
reference's outputs, so the tests need neither networkx-1.x nor the reference.
"""

import hashlib
import os
import sys

import numpy as np
import networkx as nx

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from pypercolate_b200 import lowering  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def lower_ref(pg):
    """Lower a reference ``percolation_graph`` dict with the product's lowering
    (the fixtures pin that lowering against the reference's outputs)."""
    if pg['spanning_cluster']:
        return lowering.lower(pg['perc_graph'], True, pg['auxiliary_node_attributes'],
                              pg['auxiliary_edge_attributes'], pg['spanning_sides'])
    return lowering.lower(pg['perc_graph'], False)


def graph_arrays(g):
    d = dict(N=g.num_nodes, M=g.num_edges, eu=g.eu, ev=g.ev,
             preconnected=int(g.preconnected))
    d['side_mask'] = g.side_mask if g.side_mask is not None else np.zeros(0, np.uint8)
    d['spanning'] = int(g.side_mask is not None)
    return d


def rows_bytes(rows):
    """Packed rows as uint8 with the undefined ``edge[0]`` zeroed."""
    r = rows.copy()
    r['edge'][0] = 0
    return r.view(np.uint8).copy()


def kat_graph(span):
    """The fixture of percolate/test/test_percolate.py:49-73."""
    ret = nx.Graph()
    ret.add_nodes_from(range(9))
    ret.add_edges_from([(i, i + j) for i in [1, 4, 7] for j in [-1, 1]])
    ret.add_edges_from([(i, i + j) for i in [3, 4, 5] for j in [-3, 3]])
    if span:
        ret.add_nodes_from(range(9, 12), span=0)
        ret.add_nodes_from(range(12, 15), span=1)
        ret.add_edges_from([(0, 9), (3, 10), (6, 11)], span=0)
        ret.add_edges_from([(2, 12), (5, 13), (8, 14)], span=1)
    return ret


def odd_graph():
    """A non-lattice graph: random bonds, uneven auxiliary structure, one real
    node touching both sides' auxiliary nodes is avoided (see preconnected)."""
    g = nx.gnm_random_graph(40, 70, seed=3)
    g.add_node('L0', span=0)
    g.add_node('L1', span=0)
    g.add_node('R0', span=1)
    g.add_edges_from([('L0', 0), ('L0', 5), ('L1', 7)], span=0)
    g.add_edges_from([('R0', 33), ('R0', 39)], span=1)
    g.add_edge('L0', 'L1')          # plain edge between auxiliary nodes
    return g


def preconnected_graph():
    """Real node 2 carries auxiliary edges of BOTH sides: the sides are joined
    before any bond is added, the flag turns True at the first merge
    (percolate/hpc.py:266-274)."""
    g = nx.path_graph(6)
    g.add_node('a', span=0)
    g.add_node('b', span=1)
    g.add_edge('a', 2, span=0)
    g.add_edge('b', 2, span=1)
    return g


def hpc_fixture(name, graph, spanning, seeds, p, h, ps=None, alpha=None):
    pg = p.percolation_graph(graph, spanning_cluster=spanning)
    low = lower_ref(pg)
    d = graph_arrays(low)
    d['seeds'] = np.asarray(seeds, dtype=np.uint64)
    rows = [h.bond_microcanonical_statistics(seed=int(s), **pg) for s in seeds]
    d['rows'] = np.stack([rows_bytes(r) for r in rows])
    d['perms'] = np.stack([r['edge'][1:].astype(np.int32) for r in rows]) \
        if low.num_edges else np.zeros((len(seeds), 0), np.int32)
    if ps is not None:
        ps = np.asarray(ps, dtype=np.float64)
        d['ps'] = ps
        pmf = np.stack([p._binomial_pmf(low.num_edges, q) for q in ps])
        d['pmf'] = pmf
        cols = 7 if spanning else 6
        canon = np.zeros((len(seeds), ps.size, cols))
        runs_avg = []
        for r, row in enumerate(rows):
            stats = np.concatenate([h.bond_canonical_statistics(row, f) for f in pmf])
            o = 0
            if spanning:
                canon[r, :, 0] = stats['percolation_probability']
                o = 1
            canon[r, :, o] = stats['max_cluster_size']
            canon[r, :, o + 1:] = stats['moments']
            runs_avg.append(h.bond_initialize_canonical_averages(stats))
        d['canon_per_run'] = canon
        import functools
        red = functools.reduce(h.bond_reduce, runs_avg)
        d['reduced'] = red.view(np.uint8).copy()
        d['reduced_itemsize'] = red.dtype.itemsize
        fin = h.finalize_canonical_averages(low.num_nodes, ps, red, alpha)
        d['finalized'] = fin.view(np.uint8).copy()
        d['alpha'] = alpha
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print("wrote", name, "N=%d M=%d seeds=%d" % (low.num_nodes, low.num_edges, len(seeds)))


def original_api_fixture(name, graph, spanning, runs, ps, seed, p, alpha=None):
    alpha = p.alpha_1sigma if alpha is None else alpha
    pg = p.percolation_graph(graph, spanning_cluster=spanning)
    low = lower_ref(pg)
    d = graph_arrays(low)
    np.random.seed(seed)
    micro = p.microcanonical_averages_arrays(p.microcanonical_averages(
        graph, runs=runs, spanning_cluster=spanning, alpha=alpha))
    # the same global stream replayed: the bond orders the runs used
    np.random.seed(seed)
    d['perms'] = np.stack([np.random.permutation(low.num_edges) for _ in range(runs)]).astype(np.int32)
    canon = p.canonical_averages(np.asarray(ps, dtype=np.float64), micro)
    d.update(runs=runs, seed=seed, alpha=alpha, ps=np.asarray(ps, dtype=np.float64))
    for k, v in micro.items():
        d['micro_' + k] = v
    for k, v in canon.items():
        d['canon_' + k] = v
    # the dict form of single states (sample_states) for the first run
    np.random.seed(seed)
    states = list(p.sample_states(graph, spanning_cluster=spanning))
    d['states_max'] = np.array([s['max_cluster_size'] for s in states], dtype=np.float64)
    d['states_moments'] = np.stack([s['moments'] for s in states])
    if spanning:
        d['states_span'] = np.array([s['has_spanning_cluster'] for s in states])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print("wrote", name, "N=%d M=%d runs=%d" % (low.num_nodes, low.num_edges, runs))


def big_run_fixture(name, L, seeds, p, h):
    """Large single runs: keep a digest and a sample of rows, not the rows."""
    graph = p.spanning_2d_grid(L)
    pg = p.percolation_graph(graph, spanning_cluster=True)
    low = lower_ref(pg)
    closed = lowering.lowered_spanning_2d_grid(L)
    assert np.array_equal(low.eu, closed.eu) and np.array_equal(low.ev, closed.ev)
    assert np.array_equal(low.side_mask, closed.side_mask)
    d = dict(L=L, N=low.num_nodes, M=low.num_edges, seeds=np.asarray(seeds, dtype=np.uint64))
    digests, samples = [], []
    for s in seeds:
        rows = h.bond_microcanonical_statistics(seed=int(s), **pg)
        b = rows_bytes(rows)
        digests.append(np.frombuffer(hashlib.sha256(b.tobytes()).digest(), dtype=np.uint8))
        samples.append(b.reshape(low.num_edges + 1, -1)[::499].copy())
    d['sha256'] = np.stack(digests)
    d['sample_rows'] = np.stack(samples)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print("wrote", name)


def original_big_fixture(name, L, seed, p):
    """One large run through the ORIGINAL api (percolate/percolate.py:103-356: float64
    moments, global numpy stream): digests of the three per-n columns."""
    graph = p.spanning_2d_grid(L)
    np.random.seed(seed)
    states = list(p.sample_states(graph, spanning_cluster=True, copy_result=True))
    mx = np.array([s['max_cluster_size'] for s in states], dtype=np.float64)
    mom = np.stack([np.asarray(s['moments'], dtype=np.float64) for s in states])   # [M+1, 5]
    span = np.array([s['has_spanning_cluster'] for s in states], dtype=np.uint8)
    d = dict(L=L, seed=seed, N=states[0]['N'], M=states[0]['M'])
    for key, arr in (('max', mx), ('moments', mom), ('span', span)):
        d['sha256_' + key] = np.frombuffer(
            hashlib.sha256(np.ascontiguousarray(arr).tobytes()).digest(), dtype=np.uint8)
    d['sample_max'] = mx[::499].copy()
    d['sample_moments'] = mom[::499].copy()
    d['sample_span'] = span[::499].copy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print("wrote", name)


def main():
    os.makedirs(OUT, exist_ok=True)
    p, h = ref_shim.load()
    only = None
    for a in sys.argv[1:]:
        if a.startswith("--only="):
            only = set(a[7:].split(","))      # regenerate just these fixtures
    if only is not None:
        jobs = {
            "big_grid256": lambda: big_run_fixture("big_grid256", 256, [3939566288], p, h),
            "orig_big_grid256": lambda: original_big_fixture("orig_big_grid256", 256, 7, p),
        }
        for name in sorted(only):
            jobs[name]()
        return
    alpha = p.alpha_1sigma
    ps9 = np.concatenate([np.linspace(0.0, 1.0, 7), [0.45, 0.5]])

    hpc_fixture("hpc_kat3x3_span", kat_graph(True), True, [42, 0, 1, 2 ** 32 - 1], p, h, ps9, alpha)
    hpc_fixture("hpc_kat3x3_nospan", kat_graph(False), False, [42, 7], p, h, ps9, alpha)
    hpc_fixture("hpc_grid8", p.spanning_2d_grid(8), True,
                [0, 42, 2 ** 32 - 1, 12345, 3939566288, 5, 6, 7, 8, 9, 10, 11], p, h, ps9, alpha)
    hpc_fixture("hpc_grid3", p.spanning_2d_grid(3), True, list(range(20)), p, h, ps9, alpha)
    hpc_fixture("hpc_chain10", p.spanning_1d_chain(10), True, [1, 2, 3], p, h, ps9, alpha)
    hpc_fixture("hpc_chain1", p.spanning_1d_chain(1), True, [1], p, h)
    hpc_fixture("hpc_odd", odd_graph(), True, [11, 12, 13, 14, 15], p, h, ps9, alpha)
    hpc_fixture("hpc_odd_nospan", odd_graph(), False, [11, 12], p, h, ps9, alpha)
    hpc_fixture("hpc_preconnected", preconnected_graph(), True, [3, 4, 5], p, h)
    hpc_fixture("hpc_grid32", p.spanning_2d_grid(32), True, [100, 101, 102], p, h,
                np.linspace(0.45, 0.55, 5), alpha)

    original_api_fixture("orig_grid6", p.spanning_2d_grid(6), True, 12, ps9, 7, p)
    original_api_fixture("orig_kat3x3_nospan", kat_graph(False), False, 5, ps9, 42, p)
    original_api_fixture("orig_chain10", p.spanning_1d_chain(10), True, 40, ps9, 3, p, alpha=0.05)
    # BASELINE config 1: tutorial sizes
    original_api_fixture("orig_config1_grid32", p.spanning_2d_grid(32), True, 40,
                         np.linspace(0.45, 0.55, 100), 0, p)

    big_run_fixture("big_grid64", 64, [42, 43], p, h)
    big_run_fixture("big_grid128", 128, [42], p, h)
    # the flagship size (BASELINE config 3), one run through each api of the reference
    big_run_fixture("big_grid256", 256, [3939566288], p, h)
    original_big_fixture("orig_big_grid256", 256, 7, p)


if __name__ == "__main__":
    main()
