"""``percolate`` -- the reference's import name, resolved to the B200-native package.

User code written against andsor/pypercolate says ``import percolate`` and
``import percolate.hpc`` (e.g. the study script percolate/share/jugfile.py:25-26, which
then calls ``percolate.spanning_2d_grid``, ``percolate.percolate.percolation_graph``,
``percolate.percolate._binomial_pmf`` and ``percolate.hpc.bond_*``).  With this directory on
the path those imports bind the modules of ``pypercolate_b200`` themselves (the same module
objects, private helpers included), so such code runs unchanged on the GPU path.
"""
import sys as _sys

import pypercolate_b200 as _impl
from pypercolate_b200 import (  # noqa: F401  (the names percolate/__init__.py:88-97 re-exports)
    sample_states, single_run_arrays, microcanonical_averages, microcanonical_averages_arrays,
    canonical_averages, spanning_1d_chain, spanning_2d_grid, statistics,
)

for _name in ("hpc", "percolate", "lowering", "study", "multi", "site"):
    _mod = __import__("pypercolate_b200." + _name, fromlist=["_"])
    _sys.modules[__name__ + "." + _name] = _mod
    globals()[_name] = _mod

__version__ = _impl.__version__
