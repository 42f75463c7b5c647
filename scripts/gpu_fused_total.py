"""Scratch: total fused time (device stopwatch) for L=256 at several settings."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pypercolate_b200 import _native, lowering
L = int(sys.argv[1]) if len(sys.argv) > 1 else 256
R = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
g = lowering.lowered_spanning_2d_grid(L)
ctx = _native.Context(0); ctx.set_graph(g)
ctx.set_ps(np.linspace(0.45, 0.55, 100))
seeds = np.arange(R, dtype=np.uint32)
flags = _native.FUSE_MICRO | _native.FUSE_CANON
ctx.run_fused(R, _native.RNG_MODES[os.environ.get('PZ_RNG', 'philox')], seeds, flags)
ctx.reset_accumulators(); ctx.synchronize()
ctx.profile(True)
ctx.timer_start(); t0 = time.time()
ctx.run_fused(R, _native.RNG_MODES[os.environ.get('PZ_RNG', 'philox')], seeds, flags)
ms = ctx.timer_stop(); wall = time.time() - t0
print("env PIPE=%s CLAIM=%s: L=%d R=%d device %.1f ms wall %.1f ms -> %.3g bonds/s" % (
    os.environ.get("PZ_PIPELINE"), os.environ.get("PZ_CLAIM_LOG2"), L, R, ms, wall * 1e3, R * g.num_edges / (ms * 1e-3)))
print("   ", {k: round(v[0], 1) for k, v in ctx.profile_read().items() if v[1]})
