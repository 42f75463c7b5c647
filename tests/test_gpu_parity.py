"""Parity of the CUDA path (through the C-ABI / the public Python API) with the
reference: golden vectors produced by the unmodified reference, the CPU
oracle on seeded inputs, and size-independent properties at full size.

Bars: integer work (rows, sums, permutations) bit-exact; floating point within
1e-10 relative (RTOL) unless a looser bound is stated next to the assertion
with its reason."""
import functools
import hashlib
import os

import numpy as np
import pytest

from conftest import (ATOL, HPC_FIXTURES, ORIG_FIXTURES, RTOL, assert_rows_equal, golden_graph,
                      golden_rows, load_golden, row_dtype)

pytestmark = pytest.mark.gpu


def _native():
    from pypercolate_b200 import _native
    return _native


def ctx_for(g, force=None):
    n = _native()
    if force is None:
        os.environ.pop("PZ_FORCE_STORE", None)
    else:
        os.environ["PZ_FORCE_STORE"] = str(force)
    try:
        ctx = n.Context(0)
    finally:
        os.environ.pop("PZ_FORCE_STORE", None)
    ctx.set_graph(g)
    return ctx


# ---------------------------------------------------------------------------
# rows: bond_sample_states / bond_microcanonical_statistics
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("force", [None, 1, 2])
@pytest.mark.parametrize("name", HPC_FIXTURES)
def test_rows_match_reference_golden(name, force):
    n = _native()
    d = load_golden(name)
    g = golden_graph(d)
    ctx = ctx_for(g, force)
    ref = golden_rows(d)
    # exact mode: the reference's own bond orders, supplied by the host
    rows = ctx.run_rows(ref.shape[0], n.PERM_HOST, d['perms'])
    assert_rows_equal(rows, ref, name + " host perms")
    # seeds only: numpy's legacy stream reproduced on the device
    rows, perms = ctx.run_rows(ref.shape[0], n.PERM_MT19937, d['seeds'].astype(np.uint32),
                               want_perms=True)
    assert np.array_equal(perms, d['perms'])
    assert_rows_equal(rows, ref, name + " device mt19937")
    assert ctx.launch_count > 0
    ctx.close()


@pytest.mark.parametrize("name", ["big_grid64", "big_grid128", "big_grid256"])
def test_large_single_runs_match_reference_digest(name):
    from pypercolate_b200 import lowering
    n = _native()
    d = load_golden(name)
    g = lowering.lowered_spanning_2d_grid(int(d['L']))
    ctx = ctx_for(g)
    rows = ctx.run_rows(len(d['seeds']), n.PERM_MT19937, d['seeds'].astype(np.uint32))
    for r in range(len(d['seeds'])):
        b = rows[r].copy()
        b['edge'][0] = 0
        raw = b.view(np.uint8)
        assert np.array_equal(raw.reshape(g.num_edges + 1, -1)[::499], d['sample_rows'][r])
        assert hashlib.sha256(raw.tobytes()).digest() == d['sha256'][r].tobytes()
    ctx.close()


@pytest.mark.parametrize("force", [None, 2])
@pytest.mark.parametrize("kind,L,runs", [("2d", 32, 48), ("2d", 128, 40), ("2d", 181, 6),
                                         ("2d", 256, 6), ("3d", 20, 8), ("chain", 200, 5)])
def test_rows_match_oracle(kind, L, runs, force):
    from pypercolate_b200 import lowering
    from oracle import oracle
    n = _native()
    g = {"2d": lowering.lowered_spanning_2d_grid, "3d": lowering.lowered_spanning_3d_grid,
         "chain": lowering.lowered_spanning_1d_chain}[kind](L)
    ctx = ctx_for(g, force)
    seeds = np.arange(runs, dtype=np.uint32) * 7919 + 13
    rows = ctx.run_rows(runs, n.PERM_MT19937, seeds)
    for r in range(runs):
        ref = oracle.sweep_rows(g.num_nodes, g.num_edges, g.eu, g.ev, g.side_mask,
                                g.preconnected, oracle.numpy_permutation(int(seeds[r]), g.num_edges))
        assert_rows_equal(rows[r], ref, "%s L=%d run %d" % (kind, L, r))
    ctx.close()


def test_rows_without_spanning_and_ragged_inputs():
    from pypercolate_b200 import lowering
    from oracle import oracle
    n = _native()
    # self loops, multi-edges, isolated nodes, empty bond list, zero runs
    g = lowering.LoweredGraph(7, [0, 1, 1, 2, 3, 3], [1, 0, 1, 3, 2, 4])
    ctx = ctx_for(g)
    perms = np.stack([np.random.RandomState(s).permutation(6) for s in range(9)]).astype(np.int32)
    rows = ctx.run_rows(9, n.PERM_HOST, perms)
    assert rows.dtype.itemsize == 52
    for r in range(9):
        assert_rows_equal(rows[r], oracle.sweep_rows(7, 6, g.eu, g.ev, None, False, perms[r]))
    assert ctx.run_rows(0, n.PERM_HOST, np.zeros((0, 6), np.int32)).shape == (0, 7)
    ctx.close()
    g0 = lowering.LoweredGraph(3, [], [], side_mask=[1, 0, 2])
    ctx = ctx_for(g0)
    rows = ctx.run_rows(2, n.PERM_HOST, np.zeros((2, 0), np.int32))
    assert rows.shape == (2, 1) and rows['max_cluster_size'].tolist() == [[1], [1]]
    assert rows['moments'][0, 0].tolist() == [2] * 5
    ctx.close()


def test_philox_bond_orders_match_restatement():
    from pypercolate_b200 import lowering
    from oracle import oracle
    n = _native()
    # (up to L = 256 the bucket count stops at 2048; beyond, it grows with the number of bonds)
    for L in (2, 3, 8, 32, 100, 128, 181, 256, 300, 512, 1024):
        g = lowering.lowered_spanning_2d_grid(L)
        ctx = ctx_for(g)
        seeds = np.array([0, 1, 42, 2 ** 32 - 1, 99], dtype=np.uint32)
        perms = ctx.make_perms(seeds.size, n.PERM_PHILOX, seeds)
        for r, s in enumerate(seeds):
            assert np.array_equal(perms[r], oracle.philox_permutation(int(s), g.num_edges))
        ctx.close()


# ---------------------------------------------------------------------------
# fused reduction over runs
# ---------------------------------------------------------------------------
def exact_sums(rows, N):
    mx = np.stack([r['max_cluster_size'] for r in rows]).astype(object)
    mom = np.stack([r['moments'] for r in rows]).astype(object)
    return mx, mom


def acc_totals(acc):
    """Exact sums behind the limb words of a micro accumulator block (the split of a
    total into limb words depends on how the runs were grouped; the totals do not)."""
    a = acc.astype(object)
    cols = [a[:, 0], a[:, 1], a[:, 2] + (a[:, 3] << 32), a[:, 4], a[:, 5] + (a[:, 6] << 32)]
    for k in range(3):
        q = a[:, 7 + 6 * k: 13 + 6 * k]
        cols.append(q[:, 0] + (q[:, 1] << 32))
        cols.append(q[:, 2] + (q[:, 3] << 32) + (q[:, 4] << 64) + (q[:, 5] << 96))
    return np.stack(cols, axis=1)


@pytest.mark.parametrize("force", [None, 2])
@pytest.mark.parametrize("kind,L,runs", [("2d", 8, 70), ("2d", 32, 40), ("3d", 6, 33), ("2d", 40, 300)])
def test_micro_accumulators_are_exact(kind, L, runs, force):
    from pypercolate_b200 import lowering
    from oracle import oracle
    n = _native()
    g = (lowering.lowered_spanning_2d_grid if kind == "2d" else lowering.lowered_spanning_3d_grid)(L)
    N, M = g.num_nodes, g.num_edges
    ctx = ctx_for(g, force)
    perms = np.stack([oracle.numpy_permutation(500 + r, M) for r in range(runs)])
    rows = [oracle.sweep_rows(N, M, g.eu, g.ev, g.side_mask, g.preconnected, p) for p in perms]
    # two calls accumulate
    half = runs // 2
    ctx.run_fused(half, n.PERM_HOST, perms[:half], n.FUSE_MICRO)
    ctx.run_fused(runs - half, n.PERM_HOST, perms[half:], n.FUSE_MICRO)
    assert ctx.micro_runs == runs
    acc = ctx.micro_export()
    mx, mom = exact_sums(rows, N)
    span = np.stack([r['has_spanning_cluster'] for r in rows]).astype(np.int64)
    assert np.array_equal(np.cumsum(acc[:, 0].astype(np.int64)), span.sum(axis=0))
    step = max(1, (M + 1) // 200)
    for i in list(range(0, M + 1, step)) + [M]:
        w = [int(x) for x in acc[i]]
        assert w[1] == sum(mx[:, i])
        assert w[2] + (w[3] << 32) == sum(x * x for x in mx[:, i])
        c = [(N - 1) - x for x in mom[:, i, 0]]
        assert w[4] == sum(c) and w[5] + (w[6] << 32) == sum(x * x for x in c)
        for k in range(3):
            q = w[7 + 6 * k: 13 + 6 * k]
            assert q[0] + (q[1] << 32) == sum(mom[:, i, 2 + k])
            assert q[2] + (q[3] << 32) + (q[4] << 64) + (q[5] << 96) == \
                sum(x * x for x in mom[:, i, 2 + k])
    mean, var = ctx.micro_finalize()
    fm = np.stack([r['max_cluster_size'] for r in rows]).astype(np.float64)
    fmom = np.stack([r['moments'] for r in rows]).astype(np.float64)
    assert np.array_equal(mean[0], span.sum(axis=0))
    np.testing.assert_allclose(mean[1], fm.mean(axis=0), rtol=1e-13)
    np.testing.assert_allclose(var[0], fm.var(axis=0, ddof=1), rtol=RTOL, atol=ATOL)
    for k in range(5):
        np.testing.assert_allclose(mean[2 + k], fmom[:, :, k].mean(axis=0), rtol=1e-13)
        # The device value is the exact rational rounded once (tests/test_gpu_round2.py checks it
        # against fractions); numpy's two-pass variance of the float64 data carries an error of
        # about eps * mean^2 / var of its own, so that is what bounds the comparison -- 1e-10 where
        # the data allow it
        want = fmom[:, :, k].var(axis=0, ddof=1)
        m_ = fmom[:, :, k].mean(axis=0)
        with np.errstate(divide='ignore', invalid='ignore'):
            allowed = np.where(want > 0, RTOL + 32 * np.finfo(float).eps * (m_ * m_ + want) / want, 0.0)
        assert np.all(np.abs(var[1 + k] - want) <= allowed * want + ATOL), k
        assert np.array_equal(var[1 + k] == 0, np.ptp(fmom[:, :, k], axis=0) == 0)
    # export / import round trip (the cross-GPU exchange)
    ctx.reset_accumulators()
    assert ctx.micro_runs == 0
    ctx.micro_import(acc, runs)
    assert np.array_equal(ctx.micro_export(), acc)
    ctx.close()


# ---------------------------------------------------------------------------
# canonical ensemble
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("name", [n for n in HPC_FIXTURES if n not in ("hpc_chain1", "hpc_preconnected")])
def test_canonical_chain_matches_reference_golden(name):
    from pypercolate_b200 import hpc
    n = _native()
    d = load_golden(name)
    spanning = bool(int(d['spanning']))
    g = golden_graph(d)
    ctx = ctx_for(g)
    ps = d['ps']
    pmf = ctx.set_ps(ps, want_pmf=True)
    np.testing.assert_allclose(pmf, d['pmf'], rtol=RTOL, atol=ATOL)
    runs = d['perms'].shape[0]
    ctx.run_fused(runs, n.PERM_HOST, d['perms'], n.FUSE_CANON)
    per = ctx.canon_last_runs(runs)
    ref = d['canon_per_run']
    if not spanning:
        assert np.all(per[:, :, 0] == 0)
        per = per[:, :, 1:]
    np.testing.assert_allclose(per, ref, rtol=RTOL, atol=ATOL)
    # host rows -> bond_canonical_statistics (hpc.py:443-515)
    rows = golden_rows(d)
    for i in (0, len(ps) // 2, len(ps) - 1):
        st = hpc.bond_canonical_statistics(rows[0].astype(np.dtype(hpc.microcanonical_statistics_dtype(spanning))),
                                           d['pmf'][i])
        o = 1 if spanning else 0
        if spanning:
            np.testing.assert_allclose(st['percolation_probability'][0], ref[0, i, 0], rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(st['max_cluster_size'][0], ref[0, i, o], rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(st['moments'][0], ref[0, i, o + 1:], rtol=RTOL, atol=ATOL)
    # device reduction == functools.reduce(bond_reduce, ...) of the reference
    count, mean, m2 = ctx.canon_export()
    red = hpc._canonical_averages_from_partials(count, mean, m2, spanning)
    ref_red = d['reduced'].view(np.dtype(hpc.canonical_averages_dtype(spanning)))
    assert (red['number_of_runs'] == runs).all()
    for f in red.dtype.names:
        if f.endswith('_mean'):
            np.testing.assert_allclose(red[f], ref_red[f], rtol=RTOL, atol=ATOL)
        elif f.endswith('_m2'):
            # M2 is a sum of squared differences: absolute accuracy scales with mean^2
            scale = np.abs(ref_red[f.replace('_m2', '_mean')]).max() ** 2
            # (relative bar 1e-10; the absolute term covers columns whose runs agree to many digits,
            # where the reference's own M2 is rounding noise: measured in tests/test_gpu_round2.py)
            np.testing.assert_allclose(red[f], ref_red[f], rtol=RTOL, atol=1e-13 * scale)
    fin = hpc.finalize_canonical_averages(g.num_nodes, ps, red, float(d['alpha']))
    ref_fin = d['finalized'].view(np.dtype(hpc.finalized_canonical_averages_dtype(spanning)))
    for f in fin.dtype.names:
        # std / ci inherit the absolute accuracy of sqrt(M2): scale with the mean
        base = f.rsplit('_', 1)[0] + '_mean'
        scale = np.nanmax(np.abs(ref_fin[base])) if base in ref_fin.dtype.names else 0.0
        got, want = fin[f].copy(), ref_fin[f].copy()
        if f.endswith('_ci'):
            # scipy returns NaN bounds when the runs agree to the last bit (std == 0); whether
            # they do depends on summation order, so a NaN bound is equivalent to the mean
            mean_g = np.broadcast_to(fin[base][..., None], got.shape)
            mean_w = np.broadcast_to(ref_fin[base][..., None], want.shape)
            got = np.where(np.isnan(got), mean_g, got)
            want = np.where(np.isnan(want), mean_w, want)
        np.testing.assert_allclose(got, want, rtol=RTOL, atol=1e-12 * scale, equal_nan=True)
    ctx.close()


def test_binomial_pmf_drop_in():
    import scipy.stats
    from pypercolate_b200 import percolate
    from oracle import oracle
    # percolate/test/test_percolate.py:283-295
    np.testing.assert_allclose(percolate._binomial_pmf(1000, 0.01).sum(), 1.0)
    np.testing.assert_allclose(percolate._binomial_pmf(100, 0.1),
                               scipy.stats.binom.pmf(np.arange(101), n=100, p=0.1))
    for M, p in [(12, 0.0), (12, 1.0), (1984, 0.5), (130560, 0.45), (130560, 0.5), (130560, 0.999),
                 (2095104, 0.5)]:
        got = percolate._binomial_pmf(M, p)
        np.testing.assert_allclose(got, oracle.binomial_pmf(M, p), rtol=1e-12, atol=0)


# ---------------------------------------------------------------------------
# the public Python API (drop-in behaviour)
# ---------------------------------------------------------------------------
def kat_graph(span):
    import networkx as nx
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


@pytest.mark.parametrize("span", [True, False])
def test_sample_states_known_answers(span):
    # percolate/test/test_percolate.py:92-192
    from pypercolate_b200 import percolate
    from test_oracle import KAT_EDGES, KAT_MAX, KAT_MOMENTS, KAT_PERM, KAT_SPAN
    it = percolate.sample_states(kat_graph(span), spanning_cluster=span)
    first = next(it)
    assert first['n'] == 0 and first['N'] == 9 and first['M'] == 12
    assert first['max_cluster_size'] == 1 and np.array_equal(first['moments'], np.ones(5) * 8)
    assert ('has_spanning_cluster' in first) == span and 'edge' not in first
    np.random.seed(42)            # drawn only when advanced past n == 0
    states = [first] + list(it)
    assert [s['n'] for s in states] == list(range(13))
    assert [s['max_cluster_size'] for s in states] == KAT_MAX
    for s in states:
        assert np.array_equal(s['moments'], KAT_MOMENTS[s['n']])
        assert s['moments'].dtype == np.float64
    assert [s['edge'] for s in states[1:]] == [KAT_EDGES[e] for e in KAT_PERM]
    if span:
        assert [s['has_spanning_cluster'] for s in states] == KAT_SPAN
    # copy_result=False hands out one dict
    np.random.seed(42)
    it = percolate.sample_states(kat_graph(span), spanning_cluster=span, copy_result=False)
    objs = {id(s) for s in it}
    assert len(objs) == 1


def test_bond_sample_states_and_statistics_api():
    # percolate/test/test_hpc.py:250-334
    from pypercolate_b200 import hpc, percolate
    d = load_golden("hpc_grid3")
    ref = golden_rows(d)
    pg = percolate.percolation_graph(percolate.spanning_2d_grid(3))
    for r, seed in enumerate(d['seeds'][:5]):
        arr = hpc.bond_microcanonical_statistics(seed=int(seed), **pg)
        assert arr.dtype == np.dtype(hpc.microcanonical_statistics_dtype(True))
        assert_rows_equal(arr, ref[r].astype(arr.dtype))
        gen = hpc.bond_sample_states(seed=int(seed), **pg)
        objs = set()
        for n, state in enumerate(gen):
            objs.add(id(state))
            assert state.shape == (1,) and state['n'][0] == n
            if n:
                assert state['edge'][0] == arr['edge'][n]
            assert state['max_cluster_size'][0] == arr['max_cluster_size'][n]
            assert np.array_equal(state['moments'][0], arr['moments'][n])
        assert n == pg['num_edges'] and len(objs) == 1
    # non-integer seeds go through RandomState on the host (hpc.py:195)
    arr = hpc.bond_microcanonical_statistics(seed=[1, 2, 3], **pg)
    perm = np.random.RandomState([1, 2, 3]).permutation(12)
    assert np.array_equal(arr['edge'][1:], perm)
    batch = hpc.bond_microcanonical_statistics_batch(seeds=d['seeds'].astype(np.uint32), **pg)
    assert_rows_equal(batch, ref.astype(batch.dtype))


@pytest.mark.parametrize("name", ORIG_FIXTURES)
def test_original_api_matches_reference_golden(name):
    """np.random.seed(s); statistics(...) draws the same bond orders as the
    reference and must reproduce its microcanonical and canonical averages."""
    from pypercolate_b200 import percolate
    d = load_golden(name)
    g = golden_graph(d)
    spanning = bool(int(d['spanning']))
    runs, alpha, seed = int(d['runs']), float(d['alpha']), int(d['seed'])
    np.random.seed(seed)
    it = percolate.microcanonical_averages(g, runs=runs, spanning_cluster=spanning, alpha=alpha)
    micro = percolate.microcanonical_averages_arrays(it)
    assert micro['N'] == g.num_nodes and micro['M'] == g.num_edges
    keys = [k[6:] for k in d if k.startswith('micro_') and k not in ('micro_N', 'micro_M')]
    assert sorted(keys) == sorted(k for k in micro if k not in ('N', 'M'))
    for k in keys:
        scale = np.abs(d['micro_' + k]).max()
        np.testing.assert_allclose(micro[k], d['micro_' + k], rtol=RTOL, atol=1e-13 * scale, err_msg=k)
    canon = percolate.canonical_averages(d['ps'], micro)
    for k in (k[6:] for k in d if k.startswith('canon_')):
        if k in ('N', 'M', 'ps'):
            assert np.array_equal(canon[k], d['canon_' + k])
            continue
        scale = np.abs(d['canon_' + k]).max()
        np.testing.assert_allclose(canon[k], d['canon_' + k], rtol=RTOL, atol=1e-13 * scale, err_msg=k)
    # generator form: per-n dictionaries, n == 0 before any random draw
    np.random.seed(seed)
    dicts = list(percolate.microcanonical_averages(g, runs=runs, spanning_cluster=spanning, alpha=alpha))
    assert [x['n'] for x in dicts] == list(range(g.num_edges + 1))
    via = percolate.microcanonical_averages_arrays(iter(dicts))
    for k in keys:
        np.testing.assert_array_equal(via[k], micro[k])
    # statistics() end to end
    np.random.seed(seed)
    stats = percolate.statistics(g, d['ps'], spanning_cluster=spanning, alpha=alpha, runs=runs)
    for k in canon:
        np.testing.assert_array_equal(stats[k], canon[k])
    # sample_states of the first run
    np.random.seed(seed)
    states = list(percolate.sample_states(g, spanning_cluster=spanning))
    assert [s['max_cluster_size'] for s in states] == d['states_max'].tolist()
    np.testing.assert_allclose(np.stack([s['moments'] for s in states]), d['states_moments'], rtol=RTOL)
    single = None
    np.random.seed(seed)
    single = percolate.single_run_arrays(graph=g, spanning_cluster=spanning)
    np.testing.assert_array_equal(single['max_cluster_size'], d['states_max'])
    assert single['moments'].shape == (5, g.num_edges + 1)
    if spanning:
        assert np.array_equal(single['has_spanning_cluster'], d['states_span'])


def test_original_api_large_run_matches_reference_digest():
    """One L = 256 run through ``single_run_arrays`` (global numpy stream, float64 moments) against
    the digests of the unmodified reference's ``sample_states`` (tests/golden/orig_big_grid256)."""
    from pypercolate_b200 import lowering, percolate
    d = load_golden("orig_big_grid256")
    g = lowering.lowered_spanning_2d_grid(int(d['L']))
    np.random.seed(int(d['seed']))
    single = percolate.single_run_arrays(graph=g, spanning_cluster=True)
    assert (single['N'], single['M']) == (int(d['N']), int(d['M']))
    cols = {'max': single['max_cluster_size'],
            'moments': np.ascontiguousarray(single['moments'].T),
            'span': single['has_spanning_cluster'].astype(np.uint8)}
    for key, arr in cols.items():
        assert arr.dtype == d['sample_' + key].dtype, key
        assert np.array_equal(arr[::499], d['sample_' + key]), key
        assert hashlib.sha256(arr.tobytes()).digest() == d['sha256_' + key].tobytes(), key


def test_microcanonical_averages_initial_iteration():
    # percolate/test/test_percolate.py:234-265
    import scipy.stats
    from pypercolate_b200 import percolate
    for span in (True, False):
        state = np.random.get_state()[1].copy()
        ret = next(percolate.microcanonical_averages(kat_graph(span), spanning_cluster=span))
        assert np.array_equal(np.random.get_state()[1], state)     # nothing drawn yet
        assert ret['n'] == 0 and ret['max_cluster_size'] == 1.0
        np.testing.assert_allclose(ret['max_cluster_size_ci'], np.ones(2))
        np.testing.assert_allclose(ret['moments'], np.ones(5) * 8)
        np.testing.assert_allclose(ret['moments_ci'], np.ones((5, 2)) * 8)
        assert ('spanning_cluster' in ret) == span
        if span:
            assert ret['spanning_cluster'] == 1 / 42
            np.testing.assert_allclose(
                np.array([0, 1]) + np.array([1, -1]) *
                scipy.stats.beta.cdf(ret['spanning_cluster_ci'], a=1, b=41),
                scipy.stats.norm.cdf(-1) * np.ones(2))


def test_fused_hpc_batch_equals_reference_map_reduce():
    """bond_canonical_averages_batch == reduce(bond_reduce, map(bond_run, seeds))
    (percolate/share/jugfile.py:57-135), checked against the reference's
    per-run values from the golden file."""
    from pypercolate_b200 import hpc, percolate
    d = load_golden("hpc_grid8")
    pg = percolate.percolation_graph(percolate.spanning_2d_grid(8))
    got = hpc.bond_canonical_averages_batch(seeds=d['seeds'].astype(np.uint32), ps=d['ps'], **pg)
    ref = d['reduced'].view(np.dtype(hpc.canonical_averages_dtype(True)))
    for f in got.dtype.names:
        scale = np.abs(ref[f.replace('_m2', '_mean')]).max() ** (2 if f.endswith('_m2') else 1)
        np.testing.assert_allclose(got[f], ref[f], rtol=RTOL, atol=1e-13 * scale)


def test_device_micro_arrays_equal_host_evaluation():
    """pz_micro_arrays (means, Student-t intervals, per-site normalisation on the device) is
    bit-identical to the host formulas of percolate._arrays_from_device + the division by N
    (percolate/percolate.py:613-635, 681-705, 1056-1060), including the zero-variance rows."""
    import scipy.stats
    from pypercolate_b200 import lowering, percolate
    n = _native()
    for L, runs, alpha in ((8, 70, 0.05), (32, 40, percolate.alpha_1sigma), (5, 1, 0.3), (-10, 30, 0.1)):
        # (2D grids have an odd number of rows M + 1; the chain of 10 has an even one)
        g = lowering.lowered_spanning_2d_grid(L) if L > 0 else lowering.lowered_spanning_1d_chain(-L)
        ctx = ctx_for(g)
        ctx.reset_accumulators()
        ctx.run_fused(runs, n.PERM_MT19937, np.arange(runs, dtype=np.uint32) + 17, n.FUSE_MICRO)
        mean, var = ctx.micro_finalize()
        host = percolate._arrays_from_device(mean, var, runs, alpha, g.num_nodes, g.num_edges, True)
        with np.errstate(invalid='ignore'):
            t_lo, t_hi = scipy.stats.t.interval(1 - alpha, df=runs - 1)
        k, mx, mx_ci, mom, mom_ci = ctx.micro_arrays(t_lo, t_hi, norm=g.num_nodes)
        assert np.array_equal(k, mean[0])
        for dev, key in ((mx, 'max_cluster_size'), (mx_ci, 'max_cluster_size_ci'),
                         (mom, 'moments'), (mom_ci, 'moments_ci')):
            want = host[key] / g.num_nodes
            assert dev.shape == want.shape, key
            assert np.array_equal(dev, want, equal_nan=True), (L, runs, key)
        if runs > 1:
            assert (var[0] == 0).any() and np.array_equal(mx_ci[var[0] == 0, 0], mx[var[0] == 0])
        _, mx1, mx1_ci, _, _ = ctx.micro_arrays(t_lo, t_hi)              # norm = 1
        assert np.array_equal(mx1, mean[1]) and np.array_equal(mx1_ci, host['max_cluster_size_ci'],
                                                               equal_nan=True)
        ctx.reset_accumulators()
        ctx.close()


def test_study_driver_equals_jugfile_pipeline(tmp_path):
    """finite_size_study == the jugfile's task graph (percolate/share/jugfile.py:166-259):
    same seeds recipe, reduce(bond_reduce, map(bond_run, seeds)), finalize -- checked
    against the oracle's restatement of that chain, and the on-disk format."""
    from pypercolate_b200 import study, lowering, hpc
    from oracle import oracle
    dims, runs = (4, 8), 12
    ps = np.linspace(0.4, 0.6, 5)
    out = str(tmp_path / "study.npz")
    got = study.finite_size_study(dims, number_of_runs=runs, ps=ps, output=out)
    master = np.random.RandomState(seed=study.DEFAULT_SEED)
    for L in dims:
        seeds = master.randint(study.UINT32_MAX, size=runs)
        g = lowering.lowered_spanning_2d_grid(L)
        pmfs = [oracle.binomial_pmf(g.num_edges, p) for p in ps]
        reduced = None
        for s in seeds:
            rows = oracle.sweep_rows(g.num_nodes, g.num_edges, g.eu, g.ev, g.side_mask, False,
                                     oracle.numpy_permutation(int(s), g.num_edges))
            stats = np.empty(ps.size, dtype=oracle.canonical_statistics_dtype(True))
            for i, f in enumerate(pmfs):
                stats[i] = oracle.bond_canonical_statistics(rows, f)
            one = oracle.bond_initialize_canonical_averages(stats)
            reduced = one if reduced is None else oracle.bond_reduce(reduced, one)
        want = oracle.finalize_canonical_averages(g.num_nodes, ps, reduced, study.ALPHA_1SIGMA)
        assert got[L].dtype == np.dtype(hpc.finalized_canonical_averages_dtype(True))
        for f in got[L].dtype.names:
            np.testing.assert_allclose(got[L][f], want[f], rtol=RTOL, atol=1e-12, err_msg=f)
    with np.load(out) as z:
        assert sorted(z.files) == ['4', '8']
        assert np.array_equal(z['8'], got[8])
    with pytest.raises(RuntimeError):
        study.write_to_disk(out, 8, got[8])


def test_feistel_bond_orders_match_restatement():
    from pypercolate_b200 import lowering
    from oracle import oracle
    n = _native()
    for g in (lowering.lowered_spanning_1d_chain(2), lowering.lowered_spanning_2d_grid(3),
              lowering.lowered_spanning_2d_grid(8), lowering.lowered_spanning_2d_grid(32),
              lowering.lowered_spanning_2d_grid(182), lowering.lowered_spanning_2d_grid(256),
              lowering.lowered_spanning_3d_grid(20)):
        ctx = ctx_for(g)
        seeds = np.array([0, 1, 42, 2 ** 32 - 1, 99], dtype=np.uint32)
        perms = ctx.make_perms(seeds.size, n.PERM_FEISTEL, seeds)
        for r, s in enumerate(seeds):
            assert np.array_equal(perms[r], oracle.feistel_permutation(int(s), g.num_edges)), (g.num_edges, s)
        ctx.close()


# ---------------------------------------------------------------------------
# full-size properties (BASELINE configs 3-5): no oracle can walk these in bulk
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("kind,L,runs,check", [("2d", 256, 64, 2), ("2d", 1024, 4, 1), ("3d", 64, 6, 1)])
def test_full_size_invariants(kind, L, runs, check):
    from pypercolate_b200 import lowering
    from oracle import oracle
    n = _native()
    g = (lowering.lowered_spanning_2d_grid if kind == "2d" else lowering.lowered_spanning_3d_grid)(L)
    N, M = g.num_nodes, g.num_edges
    ctx = ctx_for(g)
    seeds = np.arange(runs, dtype=np.uint32) + 2024
    mode = n.PERM_PHILOX
    chunk = max(1, min(runs, (1 << 30) // ((M + 1) * 53)))
    for r0 in range(0, runs, chunk):
        rows, perms = ctx.run_rows(min(chunk, runs - r0), mode, seeds[r0:r0 + chunk], want_perms=True)
        for r in range(rows.shape[0]):
            x = rows[r]
            assert np.array_equal(np.sort(perms[r]), np.arange(M))          # a permutation
            assert np.array_equal(x['edge'][1:], perms[r].astype(np.uint32))
            assert np.array_equal(x['n'], np.arange(M + 1, dtype=np.uint32))
            mx = x['max_cluster_size'].astype(np.int64)
            assert mx[0] == 1 and mx[-1] == N and np.all(np.diff(mx) >= 0)   # connected lattice
            assert np.all(np.diff(x['has_spanning_cluster'].astype(np.int8)) >= 0)
            assert x['has_spanning_cluster'][-1] and not x['has_spanning_cluster'][0]
            mom = x['moments']
            assert np.array_equal(mom[:, 1], (N - mx).astype(np.uint64))     # sum s = N
            assert np.all(np.diff(mom[:, 0].astype(np.int64)) <= 0)          # cluster count
            assert mom[0].tolist() == [N - 1] * 5 and mom[-1].tolist() == [0] * 5
            assert np.all(mom[:, 0].astype(np.int64) + 1 + np.arange(M + 1) >= N)  # >= N - n clusters
            if r0 + r < check:     # bit-exact against the oracle on the same bond order
                ref = oracle.sweep_rows(N, M, g.eu, g.ev, g.side_mask, False, perms[r])
                assert_rows_equal(x, ref, "%s L=%d" % (kind, L))
    ctx.close()


@pytest.mark.parametrize("mode_name", ["PERM_PHILOX", "PERM_FEISTEL"])
def test_statistical_agreement_of_philox_mode(mode_name):
    """Device bond orders are validated statistically: the spanning probability
    at the 2D threshold from 2000 runs must sit inside the reference
    stream's own 5-sigma binomial interval (and vice versa)."""
    from pypercolate_b200 import lowering
    n = _native()
    g = lowering.lowered_spanning_2d_grid(32)
    runs = 2000
    res = {}
    dev_mode = getattr(n, mode_name)
    for mode in (n.PERM_MT19937, dev_mode):
        ctx = ctx_for(g)
        ctx.run_fused(runs, mode, np.arange(runs, dtype=np.uint32) + 1, n.FUSE_MICRO)
        mean, var = ctx.micro_finalize()
        res[mode] = (mean, var)
        ctx.close()
    for i in (g.num_edges // 2, int(0.45 * g.num_edges), int(0.55 * g.num_edges)):
        k0, k1 = res[n.PERM_MT19937][0][0][i], res[dev_mode][0][0][i]
        p = (k0 + k1) / (2 * runs)
        sigma = np.sqrt(2 * runs * p * (1 - p)) + 1.0
        assert abs(k0 - k1) < 5 * sigma
        m0, m1 = res[n.PERM_MT19937][0][1][i], res[dev_mode][0][1][i]
        s = np.sqrt((res[n.PERM_MT19937][1][0][i] + res[dev_mode][1][0][i]) / runs)
        assert abs(m0 - m1) < 5 * s


@pytest.mark.parametrize("L,runs", [(256, 3000), (200, 2000), (182, 1500)])
def test_finder_warps_match_lock_step_kernel(L, runs):
    """One run per SM (32768 < N <= 65536): 8 finder warps walk the next batch up the forest while
    merges are in flight (release/acquire on the root flags).  Thousands of full runs must give
    exactly the sums and canonical partials of the plain lock-step kernel, and the first runs
    the oracle's rows."""
    from pypercolate_b200 import lowering
    from oracle import oracle
    n = _native()
    g = lowering.lowered_spanning_2d_grid(L)
    N, M = g.num_nodes, g.num_edges
    seeds = np.arange(runs, dtype=np.uint32) * 7919 + 13
    ps = np.linspace(0.45, 0.55, 7)
    got = {}
    for finders in ("0", "1"):
        old = os.environ.get("PZ_FINDERS")
        os.environ["PZ_FINDERS"] = finders
        try:
            ctx = n.Context(0)
        finally:
            if old is None:
                os.environ.pop("PZ_FINDERS", None)
            else:
                os.environ["PZ_FINDERS"] = old
        ctx.set_graph(g)
        ctx.set_ps(ps)
        ctx.run_fused(runs, n.PERM_FEISTEL, seeds, n.FUSE_MICRO | n.FUSE_CANON)
        got[finders] = (ctx.micro_export(), ctx.canon_export())
        if finders == "1":
            rows, perms = ctx.run_rows(3, n.PERM_FEISTEL, seeds[:3], want_perms=True)
            for r in range(3):
                ref = oracle.sweep_rows(N, M, g.eu, g.ev, g.side_mask, False, perms[r])
                assert_rows_equal(rows[r], ref, "finder warps L=%d" % L)
        ctx.close()
    assert np.array_equal(acc_totals(got["0"][0]), acc_totals(got["1"][0]))
    assert got["0"][1][0] == got["1"][1][0]
    assert np.array_equal(got["0"][1][1], got["1"][1][1]) and np.array_equal(got["0"][1][2], got["1"][1][2])


@pytest.mark.parametrize("L,runs,hubs,epoch", [(256, 2500, 2, None), (256, 1500, 1, None), (256, 1500, 3, None),
                                               (200, 1500, 2, None), (182, 1200, 3, None),
                                               (256, 600, 2, 0x1300)])
def test_carried_bonds_and_hubs_match_lock_step_kernel(L, runs, hubs, epoch):
    """sweep_fc_kernel (PZ_FINDERS=2, the default for one run per SM): pending bonds carried into the
    next batch, star merging around PZ_HUBS hubs.  Thousands of full runs must give exactly the
    sums, canonical partials and first-spanning counts of the plain lock-step kernel, and the first
    runs the oracle's rows; a short claim epoch exercises the rebase."""
    from pypercolate_b200 import lowering
    from oracle import oracle
    n = _native()
    g = lowering.lowered_spanning_2d_grid(L)
    N, M = g.num_nodes, g.num_edges
    seeds = np.arange(runs, dtype=np.uint32) * 104729 + 71
    ps = np.linspace(0.45, 0.55, 7)
    got = {}
    for finders in ("0", "2"):
        over = {"PZ_FINDERS": finders, "PZ_HUBS": str(hubs)}
        if epoch is not None and finders == "2":
            over["PZ_EPOCH_START"] = str(epoch)
        old = {k: os.environ.get(k) for k in over}
        os.environ.update(over)
        try:
            ctx = n.Context(0)
            ctx.set_graph(g)
            ctx.set_ps(ps)
            ctx.run_fused(runs, n.PERM_FEISTEL, seeds, n.FUSE_MICRO | n.FUSE_CANON)
            got[finders] = (ctx.micro_export(), ctx.canon_export())
            if finders == "2":
                rows, perms = ctx.run_rows(4, n.PERM_FEISTEL, seeds[:4], want_perms=True)
                for r in range(4):
                    ref = oracle.sweep_rows(N, M, g.eu, g.ev, g.side_mask, False, perms[r])
                    assert_rows_equal(rows[r], ref, "carried bonds L=%d hubs=%d" % (L, hubs))
            ctx.close()
        finally:
            for k, v in old.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
    assert np.array_equal(acc_totals(got["0"][0]), acc_totals(got["2"][0]))
    assert got["0"][1][0] == got["2"][1][0]
    assert np.array_equal(got["0"][1][1], got["2"][1][1]) and np.array_equal(got["0"][1][2], got["2"][1][2])


def test_finder_warps_on_sparse_and_empty_graphs():
    """The one-run-per-SM shape (32768 < N <= 65536) on graphs that are not lattices: no bonds at
    all, fewer bonds than one batch, a long chain, a random sparse graph."""
    from pypercolate_b200 import lowering
    from oracle import oracle
    n = _native()
    N = 40000
    rng = np.random.RandomState(5)
    cases = [
        ([], []),
        (list(range(0, 300)), list(range(1, 301))),
        (list(range(0, N - 1)), list(range(1, N))),
        (rng.randint(0, N, size=60000).tolist(), rng.randint(0, N, size=60000).tolist()),
    ]
    side = np.zeros(N, dtype=np.uint8)
    side[:50] = 1
    side[-50:] = 2
    for eu, ev in cases:
        keep = [(a, b) for a, b in zip(eu, ev) if a != b]
        eu, ev = [a for a, _ in keep], [b for _, b in keep]
        g = lowering.LoweredGraph(N, eu, ev, side_mask=side)
        ctx = ctx_for(g)
        M = g.num_edges
        runs = 5
        perms = np.stack([oracle.numpy_permutation(90 + r, M) for r in range(runs)]) if M else \
            np.zeros((runs, 0), np.int32)
        rows = ctx.run_rows(runs, n.PERM_HOST, perms)
        for r in range(runs):
            ref = oracle.sweep_rows(N, M, g.eu, g.ev, g.side_mask, g.preconnected, perms[r])
            assert_rows_equal(rows[r], ref, "sparse graph M=%d" % M)
        ctx.close()


def test_missing_graph_and_bad_arguments_fail_loudly():
    n = _native()
    ctx = n.Context(0)
    with pytest.raises(n.NativeError):
        ctx.run_rows(1, n.PERM_HOST, np.zeros((1, 0), np.int32))
    with pytest.raises(n.NativeError):
        n.Context(10 ** 6)
    from pypercolate_b200 import lowering
    g = lowering.lowered_spanning_2d_grid(4)
    ctx.set_graph(g)
    with pytest.raises(n.NativeError):
        ctx.run_fused(2, n.PERM_HOST, np.zeros((2, g.num_edges), np.int32), n.FUSE_CANON)  # no ps
    with pytest.raises(n.NativeError):
        ctx.micro_finalize()                                                                # no runs
    with pytest.raises(n.NativeError):
        ctx.micro_arrays(-1.0, 1.0)                                                         # no runs
    with pytest.raises(n.NativeError):
        ctx.micro_arrays(-1.0, 1.0, norm=0.0)                                               # bad divisor
    with pytest.raises(n.NativeError):
        ctx.set_ps(np.array([1.5]))
    ctx.close()


@pytest.mark.parametrize("env", [{"PZ_PIPELINE": "1"}, {"PZ_PIPELINE": "2"}, {"PZ_SWEEP_TEAM": "0"},
                                 {"PZ_CTA_WARPS": "2"}, {"PZ_CTA_WARPS": "32"},
                                 {"PZ_CTA_WARPS": "8", "PZ_CLAIM_LOG2": "8"}, {"PZ_CHUNK_BYTES": "3000000"},
                                 {"PZ_FINDERS": "0"},
                                 {"PZ_CKPT_EVERY": "1024"}])
def test_alternative_launch_shapes_give_identical_results(env):
    """Three-stream pipelining, the single-warp A/B kernel, other CTA shapes, a tiny (collision
    heavy) claim table and many small chunks must not change a single bit."""
    from pypercolate_b200 import lowering
    from oracle import oracle
    n = _native()
    g = lowering.lowered_spanning_2d_grid(48)
    N, M = g.num_nodes, g.num_edges
    runs = 300
    seeds = np.arange(runs, dtype=np.uint32) + 77
    ps = np.linspace(0.4, 0.6, 9)
    base = ctx_for(g)
    base.set_ps(ps)
    base.run_fused(runs, n.PERM_PHILOX, seeds, n.FUSE_MICRO | n.FUSE_CANON)
    want_acc = base.micro_export()
    want_canon = base.canon_export()
    want_rows = base.run_rows(6, n.PERM_PHILOX, seeds[:6])
    base.close()
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        ctx = n.Context(0)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    ctx.set_graph(g)
    ctx.set_ps(ps)
    ctx.run_fused(runs, n.PERM_PHILOX, seeds, n.FUSE_MICRO | n.FUSE_CANON)
    assert np.array_equal(acc_totals(ctx.micro_export()), acc_totals(want_acc))
    got = ctx.canon_export()
    assert got[0] == want_canon[0]
    if "PZ_CHUNK_BYTES" in env:      # chunking changes the association of the Chan merge
        np.testing.assert_allclose(got[1], want_canon[1], rtol=1e-13)
        np.testing.assert_allclose(got[2], want_canon[2], rtol=RTOL, atol=1e-9 * np.abs(want_canon[2]).max())
    else:
        assert np.array_equal(got[1], want_canon[1]) and np.array_equal(got[2], want_canon[2])
    assert_rows_equal(ctx.run_rows(6, n.PERM_PHILOX, seeds[:6]), want_rows)
    ref = oracle.sweep_rows(N, M, g.eu, g.ev, g.side_mask, False, oracle.philox_permutation(77, M))
    assert_rows_equal(want_rows[0], ref)
    ctx.close()
