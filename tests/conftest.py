import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))


def golden_graph(d):
    """LoweredGraph from the arrays stored in a fixture."""
    from pypercolate_b200 import lowering
    sm = d['side_mask'] if int(d['spanning']) else None
    return lowering.LoweredGraph(int(d['N']), d['eu'], d['ev'], sm, bool(int(d['preconnected'])))


def row_dtype(spanning):
    fields = [('n', '<u4'), ('edge', '<u4')]
    if spanning:
        fields.append(('has_spanning_cluster', '?'))
    fields += [('max_cluster_size', '<u4'), ('moments', '<u8', (5,))]
    return np.dtype(fields)


def golden_rows(d):
    """(runs, M+1) structured rows stored as bytes in a fixture."""
    dt = row_dtype(bool(int(d['spanning'])))
    raw = np.ascontiguousarray(d['rows'])
    return raw.view(dt).reshape(raw.shape[0], -1)


def assert_rows_equal(a, b, msg=""):
    """Bit-exact comparison of packed rows; ``edge`` of row 0 is undefined in
    the reference (percolate/test/test_hpc.py:324-326)."""
    assert a.shape == b.shape, msg
    for name in a.dtype.names:
        x, y = a[name], b[name]
        if name == 'edge':
            x, y = x[..., 1:], y[..., 1:]
        assert np.array_equal(x, y), "%s field %s differs" % (msg, name)


HPC_FIXTURES = ["hpc_kat3x3_span", "hpc_kat3x3_nospan", "hpc_grid8", "hpc_grid3", "hpc_chain10",
                "hpc_chain1", "hpc_odd", "hpc_odd_nospan", "hpc_preconnected", "hpc_grid32"]
ORIG_FIXTURES = ["orig_grid6", "orig_kat3x3_nospan", "orig_chain10", "orig_config1_grid32"]

# float tolerance of the north star: 1e-10 relative (plus an absolute floor far
# below any physical value so exact zeros compare equal)
RTOL = 1e-10
ATOL = 1e-300


@pytest.fixture(scope="session")
def gpu_ctx_factory():
    from pypercolate_b200 import _native
    made = []

    def make(lowered):
        ctx = _native.Context(0)
        ctx.set_graph(lowered)
        made.append(ctx)
        return ctx
    yield make
    for c in made:
        c.close()
