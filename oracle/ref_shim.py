"""
TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Loads the UNMODIFIED reference modules ``percolate/percolate.py`` and
``percolate/hpc.py`` from ``/root/reference`` under an import-time
compatibility shim (the reference targets python 2/3 + networkx 1.x +
numpy 1.9 and does not import as-is under python 3.12 / networkx 3 / numpy 2).

``/root/reference`` exists only in the build container, not on the GPU box,
so this module is used exclusively by ``oracle/make_golden.py`` (to generate
``tests/golden/*.npz``) and by the ``not gpu`` tests that validate the C /
numpy restatement against the real reference (they skip when the reference
tree is absent).

What the shim provides (nothing in the reference files is edited):

* fake modules ``future`` / ``future.builtins`` re-exporting python-3 builtins
  (reference: percolate/percolate.py:13-17, percolate/hpc.py:12-14);
* fake ``simoa`` / ``simoa.stats`` with ``online_variance`` restated as the
  Chan et al. pairwise merge (sole call site: percolate/hpc.py:677-684; the
  package is not installed and not vendored -- see DESIGN.md "third-party
  arithmetic");
* ``np.float`` alias (percolate/percolate.py:561);
* the networkx-1.x spellings the reference uses: ``Graph.nodes_iter``
  (percolate/percolate.py:85,236,295; percolate/hpc.py:246), ``Graph.node``
  (percolate/percolate.py:86,237,923-925,958-962) and a list-returning
  ``Graph.edges()`` (percolate/percolate.py:244,303; percolate/hpc.py:205,255).
"""

import builtins
import importlib.util
import os
import sys
import types

import numpy as np
import scipy.stats  # noqa: F401  (import before touching np.float)
import networkx as nx

REFERENCE_ROOT = os.environ.get("PZ_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "percolate", "hpc.py"))


def online_variance(*args):
    """Chan et al. pairwise merge of (n, mean, M2) tuples.

    Restated from the algorithm the reference documents for
    simoa.stats.online_variance (docs/pypercolate-hpc.rst:66-68); called from
    percolate/hpc.py:677-684 with exactly two tuples.
    """
    n_a, mean_a, m2_a = args[0]
    for (n_b, mean_b, m2_b) in args[1:]:
        n_a_f = np.asarray(n_a, dtype=np.float64)
        n_b_f = np.asarray(n_b, dtype=np.float64)
        n = n_a_f + n_b_f
        delta = mean_b - mean_a
        mean = mean_a + delta * n_b_f / n
        m2 = m2_a + m2_b + delta * delta * n_a_f * n_b_f / n
        n_a, mean_a, m2_a = n_a + n_b, mean, m2
    return n_a, mean_a, m2_a


class _EdgesProxy(object):
    """``G.edges()`` with no arguments returns a list (networkx 1.x)."""

    def __init__(self, view):
        self._view = view

    def __call__(self, *args, **kwargs):
        res = self._view(*args, **kwargs)
        if not args and not kwargs:
            return list(res)
        return res

    def __iter__(self):
        return iter(self._view)

    def __len__(self):
        return len(self._view)

    def __contains__(self, item):
        return item in self._view

    def __getitem__(self, item):
        return self._view[item]

    def __getattr__(self, name):
        return getattr(self._view, name)


_loaded = None


def load():
    """Return ``(percolate_module, hpc_module)`` of the unmodified reference."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)

    # -- fake third-party modules ------------------------------------------
    if "future" not in sys.modules:
        future = types.ModuleType("future")
        fb = types.ModuleType("future.builtins")
        for name in ("ascii bytes chr dict filter hex input int map next oct "
                     "open pow range round str super zip").split():
            setattr(fb, name, getattr(builtins, name))
        future.builtins = fb
        sys.modules["future"] = future
        sys.modules["future.builtins"] = fb
    if "simoa" not in sys.modules:
        simoa = types.ModuleType("simoa")
        stats = types.ModuleType("simoa.stats")
        stats.online_variance = online_variance
        simoa.stats = stats
        sys.modules["simoa"] = simoa
        sys.modules["simoa.stats"] = stats

    # -- numpy / networkx 1.x spellings ------------------------------------
    if not hasattr(np, "float"):
        np.float = float
    if not hasattr(nx.Graph, "nodes_iter"):
        nx.Graph.nodes_iter = lambda self: iter(self.nodes)
    if not isinstance(nx.Graph.__dict__.get("node"), property):
        nx.Graph.node = property(lambda self: self.nodes)
    cp = nx.Graph.__dict__["edges"]
    if not isinstance(cp, property):
        func = cp.func if hasattr(cp, "func") else cp.fget
        nx.Graph.edges = property(lambda self: _EdgesProxy(func(self)))

    # -- load the two reference files unmodified ---------------------------
    pkg_name = "_pz_reference_percolate"
    pkg = types.ModuleType(pkg_name)
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, "percolate")]
    sys.modules[pkg_name] = pkg
    mods = []
    for name in ("percolate", "hpc"):
        spec = importlib.util.spec_from_file_location(
            "%s.%s" % (pkg_name, name),
            os.path.join(REFERENCE_ROOT, "percolate", name + ".py"),
        )
        mod = importlib.util.module_from_spec(spec)
        sys.modules[spec.name] = mod
        spec.loader.exec_module(mod)
        setattr(pkg, name, mod)
        mods.append(mod)
    _loaded = tuple(mods)
    return _loaded
