"""Host-side mirror of the reference interface (no GPU): dtypes, the
(count, mean, M2) chain of percolate.hpc, lazy error behaviour."""
import functools

import networkx as nx
import numpy as np
import pytest
import scipy.stats

from conftest import HPC_FIXTURES, RTOL, load_golden
from pypercolate_b200 import hpc, percolate
import pypercolate_b200


def test_public_surface_matches_reference_init():
    # percolate/__init__.py:88-97
    for name in ("sample_states", "single_run_arrays", "microcanonical_averages",
                 "microcanonical_averages_arrays", "canonical_averages", "spanning_1d_chain",
                 "spanning_2d_grid", "statistics"):
        assert callable(getattr(pypercolate_b200, name))
    assert pypercolate_b200.hpc is hpc and pypercolate_b200.percolate is percolate


# -- dtypes (percolate/test/test_hpc.py:25-141) -----------------------------------
@pytest.mark.parametrize("spanning", [True, False])
def test_microcanonical_statistics_dtype(spanning):
    dt = np.dtype(hpc.microcanonical_statistics_dtype(spanning))
    assert dt.itemsize == (53 if spanning else 52)
    assert dt['n'] == np.uint32 and dt['edge'] == np.uint32
    assert dt['max_cluster_size'] == np.uint32
    assert dt['moments'].shape == (5,) and dt['moments'].base == np.uint64
    assert ('has_spanning_cluster' in dt.names) == spanning
    assert dt.fields['moments'][1] == (13 if spanning else 12)


@pytest.mark.parametrize("spanning", [True, False])
def test_canonical_dtypes(spanning):
    dt = np.dtype(hpc.canonical_statistics_dtype(spanning))
    assert ('percolation_probability' in dt.names) == spanning
    assert dt['moments'].shape == (5,) and dt['max_cluster_size'] == np.float64
    dt = np.dtype(hpc.canonical_averages_dtype(spanning))
    assert dt['number_of_runs'] == np.uint32
    assert ('percolation_probability_mean' in dt.names) == spanning
    assert dt['moments_mean'].shape == (5,) and dt['moments_m2'].shape == (5,)
    dt = np.dtype(hpc.finalized_canonical_averages_dtype(spanning))
    for f in ('p', 'alpha', 'percolation_strength_mean', 'percolation_strength_std'):
        assert dt[f] == np.float64
    assert dt['percolation_strength_ci'].shape == (2,)
    assert dt['moments_ci'].shape == (5, 2)
    assert ('percolation_probability_ci' in dt.names) == spanning


# -- initialize / reduce / finalize against the reference's outputs -----------------
def canon_stats_from_golden(d, r):
    spanning = bool(int(d['spanning']))
    c = d['canon_per_run'][r]
    st = np.empty(c.shape[0], dtype=hpc.canonical_statistics_dtype(spanning))
    o = 0
    if spanning:
        st['percolation_probability'] = c[:, 0]
        o = 1
    st['max_cluster_size'] = c[:, o]
    st['moments'] = c[:, o + 1:]
    return st


@pytest.mark.parametrize("name", [n for n in HPC_FIXTURES if n not in
                                  ("hpc_chain1", "hpc_preconnected")])
def test_reduce_finalize_chain_matches_reference(name):
    d = load_golden(name)
    spanning = bool(int(d['spanning']))
    runs = d['canon_per_run'].shape[0]
    parts = [hpc.bond_initialize_canonical_averages(canon_stats_from_golden(d, r))
             for r in range(runs)]
    assert all((p['number_of_runs'] == 1).all() for p in parts)
    red = functools.reduce(hpc.bond_reduce, parts)
    ref_red = d['reduced'].view(np.dtype(hpc.canonical_averages_dtype(spanning)))
    assert (red['number_of_runs'] == runs).all()
    for f in red.dtype.names:
        np.testing.assert_allclose(red[f], ref_red[f], rtol=RTOL, atol=1e-12 * np.abs(ref_red[f]).max())
    fin = hpc.finalize_canonical_averages(int(d['N']), d['ps'], red, float(d['alpha']))
    ref_fin = d['finalized'].view(np.dtype(hpc.finalized_canonical_averages_dtype(spanning)))
    for f in fin.dtype.names:
        scale = np.nanmax(np.abs(ref_fin[f])) if np.isfinite(ref_fin[f]).any() else 0.0
        np.testing.assert_allclose(fin[f], ref_fin[f], rtol=RTOL, atol=1e-12 * scale,
                                   equal_nan=True, err_msg=f)


def test_bond_reduce_is_associative_and_matches_numpy():
    # percolate/test/test_hpc.py:454-524
    rng = np.random.RandomState(5)
    nump, runs = 6, 17
    stats = []
    for _ in range(runs):
        s = np.empty(nump, dtype=hpc.canonical_statistics_dtype(True))
        s['percolation_probability'] = rng.rand(nump)
        s['max_cluster_size'] = rng.rand(nump) * 100
        s['moments'] = rng.rand(nump, 5) * 1e6
        stats.append(s)
    parts = [hpc.bond_initialize_canonical_averages(s) for s in stats]
    left = functools.reduce(hpc.bond_reduce, parts)
    right = functools.reduce(lambda a, b: hpc.bond_reduce(b, a), parts)
    tree = hpc.bond_reduce(functools.reduce(hpc.bond_reduce, parts[:8]),
                           functools.reduce(hpc.bond_reduce, parts[8:]))
    for key in ('percolation_probability', 'max_cluster_size', 'moments'):
        x = np.stack([s[key] for s in stats])
        for red in (left, right, tree):
            np.testing.assert_allclose(red[key + '_mean'], x.mean(axis=0), rtol=1e-7)
            np.testing.assert_allclose(red[key + '_m2'], ((x - x.mean(axis=0)) ** 2).sum(axis=0),
                                       rtol=1e-7, atol=1e-6)
    assert (left['number_of_runs'] == runs).all()


def test_finalize_formulas():
    # percolate/test/test_hpc.py:527-694
    nump, runs, N, alpha = 4, 9, 50, 0.1
    avg = np.zeros(nump, dtype=hpc.canonical_averages_dtype(True))
    rng = np.random.RandomState(3)
    avg['number_of_runs'] = runs
    for key in ('percolation_probability', 'max_cluster_size', 'moments'):
        avg[key + '_mean'] = rng.rand(*avg[key + '_mean'].shape) * 10
        avg[key + '_m2'] = rng.rand(*avg[key + '_m2'].shape) * 5
    ps = np.linspace(0.1, 0.9, nump)
    fin = hpc.finalize_canonical_averages(N, ps, avg, alpha)
    assert (fin['p'] == ps).all() and (fin['alpha'] == alpha).all()
    np.testing.assert_allclose(fin['percolation_probability_mean'], avg['percolation_probability_mean'])
    np.testing.assert_allclose(fin['percolation_strength_mean'], avg['max_cluster_size_mean'] / N)
    np.testing.assert_allclose(fin['moments_std'], np.sqrt(avg['moments_m2'] / (runs - 1)) / N)
    lo, hi = fin['percolation_strength_ci'][:, 0], fin['percolation_strength_ci'][:, 1]
    scale = fin['percolation_strength_std'] / np.sqrt(runs)
    cdf = scipy.stats.t.cdf((hi - fin['percolation_strength_mean']) / scale, df=runs - 1)
    np.testing.assert_allclose(cdf, 1 - alpha / 2, rtol=1e-6)
    np.testing.assert_allclose(hi - fin['percolation_strength_mean'],
                               fin['percolation_strength_mean'] - lo, rtol=1e-9)


# -- lazy error behaviour (percolate/test/test_percolate.py:76-89, 206-231) ---------
def test_sample_states_errors_surface_on_first_next():
    it = percolate.sample_states(nx.Graph(), model='site')
    with pytest.raises(ValueError):
        next(it)
    it = percolate.sample_states(nx.Graph())
    with pytest.raises(ValueError):
        next(it)
    g = nx.path_graph(4)
    g.nodes[0]['span'] = 0
    it = percolate.sample_states(g)
    with pytest.raises(ValueError):
        next(it)


@pytest.mark.parametrize("kwargs", [dict(runs=0), dict(runs=-3), dict(runs='x'), dict(alpha=0.0),
                                    dict(alpha=1.0), dict(alpha='x'), dict(alpha=-0.5)])
def test_microcanonical_averages_argument_errors(kwargs):
    it = percolate.microcanonical_averages(percolate.spanning_2d_grid(3), **kwargs)
    with pytest.raises(ValueError):
        next(it)


def test_bond_sample_states_needs_two_sides():
    g = nx.path_graph(3)
    it = hpc.bond_sample_states(g, 3, 2, seed=1, spanning_cluster=True,
                                auxiliary_node_attributes={}, auxiliary_edge_attributes={},
                                spanning_sides=[0])
    with pytest.raises(ValueError):
        next(it)


def test_alpha_1sigma():
    assert percolate.alpha_1sigma == 2 * scipy.stats.norm.cdf(-1.0)


def test_spanning_graph_builders():
    g = percolate.spanning_2d_grid(3)
    assert g.number_of_nodes() == 15
    assert sum(1 for _, a in g.nodes(data=True) if 'span' in a) == 6
    assert sum(1 for *_, a in g.edges(data=True) if 'span' in a) == 6
    c = percolate.spanning_1d_chain(4)
    assert c.number_of_nodes() == 6 and c.nodes[0]['span'] == 0 and c.nodes[5]['span'] == 1


def test_study_seed_recipe_and_disk_format(tmp_path):
    """The study driver draws seeds like percolate/share/jugfile.py:36-37,172,215 and
    appends one dataset per system size, refusing to overwrite (jugfile.py:138-156)."""
    from pypercolate_b200 import study
    assert study.DEFAULT_SEED == 201508061904 % 4294967296
    seeds = study.study_seeds((8, 16, 32), 100)
    rng = np.random.RandomState(seed=study.DEFAULT_SEED)
    for L in (8, 16, 32):
        assert np.array_equal(seeds[L], rng.randint(4294967296, size=100))
    path = str(tmp_path / "out.npz")
    a = np.arange(6, dtype=np.float64).reshape(2, 3)
    study.write_to_disk(path, 8, a)
    study.write_to_disk(path, 16, a * 2)
    with np.load(path) as z:
        assert sorted(z.files) == ['16', '8'] and np.array_equal(z['16'], a * 2)
    with pytest.raises(RuntimeError):
        study.write_to_disk(path, 8, a)
    # the reference's dtypes carry np.str_ field names (percolate/hpc.py:22-31)
    b = np.zeros(3, dtype=hpc.finalized_canonical_averages_dtype(True))
    b['p'] = [0.1, 0.2, 0.3]
    study.write_to_disk(path, 32, b)
    with np.load(path) as z:
        assert z['32'].dtype == b.dtype and np.array_equal(z['32'], b)


# ---------------------------------------------------------------------------
# per-n averaging helpers (percolate/percolate.py:450-705, 968-1064)
# ---------------------------------------------------------------------------
def _random_runs(rng, runs, rows, constant_rows=()):
    """Integer per-run statistics [rows, runs] like the sweep produces; some rows identical in every run."""
    largest = rng.integers(1, 5000, size=(rows, runs)).astype(np.float64)
    moments = rng.integers(0, 2 ** 40, size=(rows, runs, 5)).astype(np.float64)
    spans = rng.random((rows, runs)) < np.linspace(0, 1, rows)[:, None]
    for n in constant_rows:
        largest[n] = largest[n, 0]
        moments[n] = moments[n, 0]
    return largest, moments, spans


def test_per_n_helpers_formulas():
    rng = np.random.default_rng(5)
    runs, alpha = 23, 0.1
    largest, moments, spans = _random_runs(rng, runs, 4, constant_rows=(2,))
    for n in range(4):
        got = percolate._microcanonical_average_max_cluster_size(largest[n], alpha)
        mean, std = largest[n].mean(), largest[n].std(ddof=1)
        assert got['max_cluster_size'] == mean
        if n == 2:                                   # percolate/percolate.py:621-633: zero spread -> (mean, mean)
            assert std == 0 and np.array_equal(got['max_cluster_size_ci'], [mean, mean])
        else:
            want = scipy.stats.t.interval(1 - alpha, df=runs - 1, loc=mean, scale=std / np.sqrt(runs))
            assert np.array_equal(got['max_cluster_size_ci'], want)
        gm = percolate._microcanonical_average_moments(moments[n], alpha)
        assert np.array_equal(gm['moments'], moments[n].mean(axis=0)) and gm['moments_ci'].shape == (5, 2)
        if n == 2:
            assert np.array_equal(gm['moments_ci'][:, 0], gm['moments']) and \
                np.array_equal(gm['moments_ci'][:, 1], gm['moments'])
        gs = percolate._microcanonical_average_spanning_cluster(spans[n], alpha)
        k = spans[n].sum()
        assert gs['spanning_cluster'] == (k + 1) / (runs + 2)
        assert np.array_equal(gs['spanning_cluster_ci'],
                              scipy.stats.beta.ppf([alpha / 2, 1 - alpha / 2], k + 1, runs - k + 1))


def test_vectorised_arrays_equal_the_per_n_helpers():
    """_arrays_from_device (the host specification of pz_micro_arrays) against the per-n functions the
    reference calls once per bond count; means and unbiased variances formed like the device does."""
    rng = np.random.default_rng(11)
    runs, rows, alpha = 40, 60, percolate.alpha_1sigma
    largest, moments, spans = _random_runs(rng, runs, rows, constant_rows=(0, 7, rows - 1))
    mean = np.empty((7, rows)); var = np.empty((6, rows))
    mean[0] = spans.sum(axis=1)
    mean[1] = largest.mean(axis=1); var[0] = largest.var(axis=1, ddof=1)
    mean[2:7] = moments.mean(axis=1).T; var[1:6] = moments.var(axis=1, ddof=1).T
    got = percolate._arrays_from_device(mean, var, runs, alpha, 1000, rows - 1, True)
    for n in range(rows):
        a = percolate._microcanonical_average_max_cluster_size(largest[n], alpha)
        b = percolate._microcanonical_average_moments(moments[n], alpha)
        c = percolate._microcanonical_average_spanning_cluster(spans[n], alpha)
        np.testing.assert_allclose(got['max_cluster_size'][n], a['max_cluster_size'], rtol=1e-13)
        np.testing.assert_allclose(got['max_cluster_size_ci'][n], a['max_cluster_size_ci'], rtol=1e-12)
        np.testing.assert_allclose(got['moments'][:, n], b['moments'], rtol=1e-13)
        np.testing.assert_allclose(got['moments_ci'][:, n], b['moments_ci'], rtol=1e-12)
        assert got['spanning_cluster'][n] == c['spanning_cluster']
        np.testing.assert_allclose(got['spanning_cluster_ci'][n], c['spanning_cluster_ci'], rtol=1e-12)
    for n in (0, 7, rows - 1):                       # identical runs: the interval collapses exactly
        assert np.array_equal(got['max_cluster_size_ci'][n], [got['max_cluster_size'][n]] * 2)


def _per_n_dicts(rng, rows, spanning):
    out = []
    for n in range(rows):
        d = {'n': n, 'N': 50, 'M': rows - 1, 'max_cluster_size': rng.random() * 50,
             'max_cluster_size_ci': rng.random(2) * 50, 'moments': rng.random(5) * 1e6,
             'moments_ci': rng.random((5, 2)) * 1e6}
        if spanning:
            d['spanning_cluster'] = rng.random()
            d['spanning_cluster_ci'] = rng.random(2)
        out.append(d)
    return out


@pytest.mark.parametrize("spanning", [True, False])
def test_microcanonical_averages_arrays_from_any_iterable(spanning):
    """The generic path (an iterable of per-n dictionaries, percolate/percolate.py:1022-1064): stacking,
    the (5, M+1[, 2]) layout of the moments and the division by N of everything but the spanning keys."""
    dicts = _per_n_dicts(np.random.default_rng(3), 9, spanning)
    got = percolate.microcanonical_averages_arrays(iter(dicts))
    assert got['M'] == 8 and got['N'] == 50
    assert ('spanning_cluster' in got) == spanning
    for n, d in enumerate(dicts):
        assert got['max_cluster_size'][n] == d['max_cluster_size'] / 50
        assert np.array_equal(got['max_cluster_size_ci'][n], d['max_cluster_size_ci'] / 50)
        assert np.array_equal(got['moments'][:, n], d['moments'] / 50)
        assert np.array_equal(got['moments_ci'][:, n], d['moments_ci'] / 50)
        if spanning:
            assert got['spanning_cluster'][n] == d['spanning_cluster']
            assert np.array_equal(got['spanning_cluster_ci'][n], d['spanning_cluster_ci'])
    assert got['moments'].shape == (5, 9) and got['moments_ci'].shape == (5, 9, 2)


def test_host_helpers_against_the_live_reference():
    """Same inputs through the unmodified reference's helpers (build container only)."""
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("reference tree not present")
    ref, _ = ref_shim.load()
    rng = np.random.default_rng(17)
    largest, moments, spans = _random_runs(rng, 12, 3, constant_rows=(1,))
    for n in range(3):
        for ours, theirs, arg in (
                (percolate._microcanonical_average_max_cluster_size,
                 ref._microcanonical_average_max_cluster_size, largest[n]),
                (percolate._microcanonical_average_moments, ref._microcanonical_average_moments, moments[n]),
                (percolate._microcanonical_average_spanning_cluster,
                 ref._microcanonical_average_spanning_cluster, spans[n])):
            a, b = ours(arg, 0.2), theirs(arg, 0.2)
            assert a.keys() == b.keys()
            for key in a:
                assert np.array_equal(np.asarray(a[key]), np.asarray(b[key])), key
    for spanning in (True, False):
        dicts = _per_n_dicts(np.random.default_rng(4), 6, spanning)
        a = percolate.microcanonical_averages_arrays(iter(dicts))
        b = ref.microcanonical_averages_arrays(iter(dicts))
        assert a.keys() == b.keys()
        for key in a:
            assert np.array_equal(np.asarray(a[key]), np.asarray(b[key])), key
    for L in (1, 2, 4):
        for ours, theirs in ((percolate.spanning_2d_grid, ref.spanning_2d_grid),
                             (percolate.spanning_1d_chain, ref.spanning_1d_chain)):
            a, b = ours(L), theirs(L)
            assert list(a.nodes(data=True)) == list(b.nodes(data=True))
            assert list(a.edges(data=True)) == list(b.edges(data=True))
            pa, pb = percolate.percolation_graph(a), ref.percolation_graph(b)
            assert pa.keys() == pb.keys()
            for key in pa:
                if key in ('graph', 'perc_graph'):
                    assert list(pa[key].edges()) == list(pb[key].edges())
                    assert list(pa[key].nodes()) == list(pb[key].nodes())
                else:
                    assert pa[key] == pb[key], key


def test_study_accepts_every_documented_graph_callable(monkeypatch):
    """finite_size_study(graph=...) takes a LoweredGraph, a networkx graph with auxiliary nodes or a
    percolation_graph dict; none of them may collide with the study's own keyword arguments."""
    from pypercolate_b200 import hpc, study, percolate, lowering
    seen = []

    def fake_batch(**kwargs):
        seen.append(kwargs)
        ps = kwargs['ps']
        out = np.zeros(ps.size, dtype=hpc.canonical_averages_dtype(kwargs['spanning_cluster']))
        out['number_of_runs'] = kwargs['seeds'].size
        return out

    monkeypatch.setattr(hpc, 'bond_canonical_averages_batch', fake_batch)
    kinds = {
        'lowered': lowering.lowered_spanning_2d_grid,
        'networkx': percolate.spanning_2d_grid,
        'dict': lambda L: percolate.percolation_graph(percolate.spanning_2d_grid(L)),
    }
    for name, fn in kinds.items():
        for spanning in (True, False):
            seen.clear()
            res = study.finite_size_study(system_dimensions=(3,), number_of_runs=5,
                                          ps=np.linspace(0.3, 0.7, 4), graph=fn,
                                          spanning_cluster=spanning)
            assert set(res) == {3} and res[3]['number_of_runs'].tolist() == [5] * 4
            kw = seen[0]
            # (like the reference's percolation_graph, a networkx graph keeps its auxiliary nodes
            # when no spanning cluster is to be detected, percolate/percolate.py:55-100)
            whole = name == 'networkx' and not spanning
            assert (kw['num_nodes'], kw['num_edges']) == ((15, 22) if whole else (9, 12)), name
            assert kw['spanning_cluster'] is spanning and 'graph' not in kw


def test_reference_import_names_resolve_to_this_package():
    """``import percolate`` / ``import percolate.hpc`` -- what user code and the reference's
    jugfile (percolate/share/jugfile.py:25-26) say -- bind this package's modules."""
    import importlib
    import pypercolate_b200
    import percolate
    import percolate.hpc
    import percolate.percolate
    assert percolate.hpc is pypercolate_b200.hpc
    assert percolate.percolate is pypercolate_b200.percolate
    assert importlib.import_module("percolate.hpc") is pypercolate_b200.hpc
    for name in ("sample_states", "single_run_arrays", "microcanonical_averages",
                 "microcanonical_averages_arrays", "canonical_averages", "spanning_1d_chain",
                 "spanning_2d_grid", "statistics"):
        assert getattr(percolate, name) is getattr(pypercolate_b200, name)
    # the names the jugfile reaches for
    assert callable(percolate.percolate.percolation_graph) and callable(percolate.percolate._binomial_pmf)
    for name in ("bond_microcanonical_statistics", "bond_canonical_statistics",
                 "bond_initialize_canonical_averages", "bond_reduce", "finalize_canonical_averages"):
        assert callable(getattr(percolate.hpc, name))


class _FakeH5File(object):
    """Stand-in for ``h5py.File`` (h5py is not installed in this image): the three members
    ``study.write_to_disk`` uses -- context manager, ``in``, ``create_dataset`` -- over a pickle."""

    def __init__(self, path, mode='a'):
        import pickle
        assert mode == 'a'
        self.path = path
        try:
            with open(path, 'rb') as f:
                self.data = pickle.load(f)
        except IOError:
            self.data = {}

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        import pickle
        with open(self.path, 'wb') as f:
            pickle.dump(self.data, f)

    def __contains__(self, key):
        return key in self.data

    def create_dataset(self, name, data):
        assert name not in self.data
        self.data[name] = np.array(data)


def test_study_hdf5_branch_with_a_stand_in_h5py(tmp_path, monkeypatch):
    """The HDF5 output of the study driver (percolate/share/jugfile.py:138-156: one dataset per
    system size, keyed by the size, an existing key is an error)."""
    import pickle
    import sys
    import types
    from pypercolate_b200 import hpc, study
    fake = types.ModuleType('h5py')
    fake.File = _FakeH5File
    monkeypatch.setitem(sys.modules, 'h5py', fake)
    path = str(tmp_path / 'study.h5')
    rows = np.zeros(3, dtype=hpc.finalized_canonical_averages_dtype(True))
    rows['p'] = [0.4, 0.5, 0.6]
    rows['number_of_runs'] = 7
    study.write_to_disk(path, 16, rows)
    study.write_to_disk(path, 32, rows)
    with pytest.raises(RuntimeError):
        study.write_to_disk(path, 16, rows)
    with open(path, 'rb') as f:
        stored = pickle.load(f)
    assert sorted(stored) == ['16', '32']
    assert stored['16'].dtype.names == tuple(str(n) for n in rows.dtype.names)
    assert stored['16']['p'].tolist() == [0.4, 0.5, 0.6] and (stored['32']['number_of_runs'] == 7).all()


def test_study_hdf5_branch_with_real_h5py(tmp_path):
    h5py = pytest.importorskip('h5py', reason="h5py is not installed in this image")
    from pypercolate_b200 import hpc, study
    path = str(tmp_path / 'study.h5')
    rows = np.zeros(2, dtype=hpc.finalized_canonical_averages_dtype(False))
    study.write_to_disk(path, 8, rows)
    with h5py.File(path, 'r') as f:
        assert list(f) == ['8'] and f['8'].shape == (2,)


def test_study_hdf5_without_h5py_says_so(tmp_path, monkeypatch):
    import sys
    from pypercolate_b200 import hpc, study
    monkeypatch.setitem(sys.modules, 'h5py', None)          # import h5py -> ImportError
    with pytest.raises(RuntimeError, match='h5py'):
        study.write_to_disk(str(tmp_path / 'x.h5'), 8,
                            np.zeros(1, dtype=hpc.finalized_canonical_averages_dtype(True)))
