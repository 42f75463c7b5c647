"""Scratch: per-phase device times of the fused path (CUDA events inside the library)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pypercolate_b200 import _native, lowering

cfg = [(int(a.split(':')[0]), int(a.split(':')[1])) for a in sys.argv[1:]] or [(128, 8192), (256, 2960)]
for L, R in cfg:
    g = lowering.lowered_spanning_2d_grid(L) if L > 0 else lowering.lowered_spanning_3d_grid(-L)
    ctx = _native.Context(0); ctx.set_graph(g)
    M = g.num_edges
    seeds = np.arange(R, dtype=np.uint32)
    ctx.set_ps(np.linspace(0.45, 0.55, 100))
    mode = _native.RNG_MODES[os.environ.get('PZ_RNG', 'philox')]
    flags = _native.FUSE_MICRO | _native.FUSE_CANON
    ctx.run_fused(min(R, 256), mode, seeds[:min(R, 256)], flags)   # warm-up
    ctx.reset_accumulators()
    ctx.profile(True)
    t0 = time.time(); ctx.run_fused(R, mode, seeds, flags); ctx.synchronize(); dt = time.time() - t0
    ph = ctx.profile_read()
    print("L=%d N=%d M=%d R=%d wall %.1f ms -> %.3g bonds/s" % (L, g.num_nodes, M, R, dt * 1e3, R * M / dt))
    for k, (ms, cnt) in ph.items():
        if cnt: print("   %-12s %9.2f ms (%d launches)  %.3g bonds/s" % (k, ms, cnt, R * M / (ms * 1e-3)))
    ctx.close()
