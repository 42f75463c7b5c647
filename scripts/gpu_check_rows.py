"""Scratch GPU check: run_rows vs the C oracle for several graphs / stores."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pypercolate_b200 import _native, lowering
from oracle import oracle

def check(g, runs, force=None, seed0=7):
    if force is not None:
        os.environ["PZ_FORCE_STORE"] = str(force)
    else:
        os.environ.pop("PZ_FORCE_STORE", None)
    ctx = _native.Context(0)
    ctx.set_graph(g)
    rng = np.random.RandomState(seed0)
    perms = np.stack([rng.permutation(g.num_edges).astype(np.int32) for _ in range(runs)])
    t0 = time.time()
    rows = ctx.run_rows(runs, _native.PERM_HOST, perms)
    dt = time.time() - t0
    bad = 0
    for r in range(runs):
        ref = oracle.sweep_rows(g.num_nodes, g.num_edges, g.eu, g.ev, g.side_mask, g.preconnected, perms[r])
        for name in ref.dtype.names:
            a, b = rows[r][name], ref[name]
            if name == 'edge':
                a, b = a[1:], b[1:]
            if not np.array_equal(a, b):
                bad += 1
                idx = np.argwhere(a != b)[0]
                print("  MISMATCH run", r, name, "first at", idx, a[tuple(idx)], b[tuple(idx)])
                break
    print("N=%d M=%d runs=%d force=%s: %s (%.3fs)" % (g.num_nodes, g.num_edges, runs, force, "OK" if not bad else "%d BAD" % bad, dt))
    ctx.close()
    return bad

bad = 0
for L in (3, 8, 32):
    g = lowering.lowered_spanning_2d_grid(L)
    for force in (None, 1, 2):
        bad += check(g, 40, force)
    bad += check(g.without_spanning(), 5, None)
bad += check(lowering.lowered_spanning_1d_chain(10), 10)
bad += check(lowering.lowered_spanning_1d_chain(1), 2)
g = lowering.lowered_spanning_2d_grid(128)
for force in (None, 1, 2):
    bad += check(g, 24, force)
g = lowering.lowered_spanning_2d_grid(256)
for force in (None, 2):
    bad += check(g, 8, force)
bad += check(lowering.lowered_spanning_3d_grid(20), 6)
print("TOTAL BAD", bad)
sys.exit(1 if bad else 0)
