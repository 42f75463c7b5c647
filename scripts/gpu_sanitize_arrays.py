"""compute-sanitizer pass over pz_micro_arrays (micro_finalize + micro_arrays kernels), odd and even M + 1."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pypercolate_b200 import _native, lowering

for g in (lowering.lowered_spanning_2d_grid(24), lowering.lowered_spanning_1d_chain(10)):     # M + 1 = 1105, 10
    ctx = _native.Context(0)
    ctx.set_graph(g)
    seeds = np.arange(40, dtype=np.uint32) + 3
    ctx.run_fused(40, _native.PERM_FEISTEL, seeds, _native.FUSE_MICRO)
    out = ctx.micro_arrays(-1.0, 1.0, norm=g.num_nodes)
    assert all(np.isfinite(a).all() for a in out)
    ctx.close()
    print("ok", g.num_nodes, g.num_edges + 1)
