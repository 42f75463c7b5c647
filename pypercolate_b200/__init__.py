"""pypercolate_b200 -- B200-native Newman-Ziff bond percolation.

Drop-in for the hot path of andsor/pypercolate: the names re-exported here are
the ones ``percolate/__init__.py:88-97`` of the reference re-exports, and
``pypercolate_b200.hpc`` mirrors ``percolate.hpc``.  The per-bond work runs in
hand-written sm_100a CUDA kernels behind the C-ABI of ``include/pz.h``; there
is no CPU fallback.
"""

from . import hpc, lowering, percolate, site, study  # noqa: F401
from .percolate import (  # noqa: F401
    sample_states,
    single_run_arrays,
    microcanonical_averages,
    microcanonical_averages_arrays,
    canonical_averages,
    spanning_1d_chain,
    spanning_2d_grid,
    statistics,
)

__version__ = "0.1.0"
