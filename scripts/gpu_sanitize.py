"""Small end-to-end pass for compute-sanitizer (memcheck): every kernel family once."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pypercolate_b200 import _native, lowering

for g, runs in ((lowering.lowered_spanning_2d_grid(24), 40), (lowering.lowered_spanning_2d_grid(256), 2),
                (lowering.lowered_spanning_3d_grid(8), 6)):
    for force in (None, 2):
        if force is not None:
            os.environ["PZ_FORCE_STORE"] = str(force)
        else:
            os.environ.pop("PZ_FORCE_STORE", None)
        if force == 2 and g.num_nodes > 20000:
            continue
        ctx = _native.Context(0)
        ctx.set_graph(g)
        ctx.set_ps(np.linspace(0.4, 0.6, 5))
        seeds = np.arange(runs, dtype=np.uint32) + 3
        for mode in (_native.PERM_FEISTEL, _native.PERM_PHILOX, _native.PERM_MT19937):
            ctx.reset_accumulators()
            ctx.run_fused(runs, mode, seeds, _native.FUSE_MICRO | _native.FUSE_CANON)
            ctx.micro_finalize()
            ctx.micro_arrays(-1.0, 1.0, norm=g.num_nodes)
            ctx.canon_export()
        ctx.run_rows(min(runs, 3), _native.PERM_FEISTEL, seeds[:3])
        ctx.close()
        print("ok", g.num_nodes, g.num_edges, force)
