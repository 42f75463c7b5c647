"""BASELINE config 4 (L = 1024, global-memory store): does keeping the parent arrays of the runs in
flight inside the L2 pay?  Sweep time of R concurrent runs for CTA shapes of 4..32 warps; R = 24
(96 MB of parent arrays: L2 resident), 148 (one CTA per SM), 888 (the shipped 6 CTAs per SM).
Usage: python scripts/gpu_c4_l2budget.py [L]   (every shape in a fresh process: knobs are read once)"""
import os, subprocess, sys
L = sys.argv[1] if len(sys.argv) > 1 else "1024"
child = r'''
import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np
from pypercolate_b200 import _native, lowering
L, R = int(sys.argv[1]), int(sys.argv[2])
g = lowering.lowered_spanning_2d_grid(L); M = g.num_edges
ctx = _native.Context(0); ctx.set_graph(g); ctx.set_ps(np.linspace(0.45, 0.55, 100))
seeds = np.arange(R, dtype=np.uint32) * 7 + 1
for rep in range(2):
    ctx.reset_accumulators(); ctx.profile(True)
    ctx.run_fused(R, _native.PERM_FEISTEL, seeds, _native.FUSE_MICRO | _native.FUSE_CANON)
    ph = ctx.profile_read(); ctx.profile(False)
ms = ph["sweep"][0]
print("sweep %8.2f ms  %.3e bonds/s  (%.2f ms per run in flight)" % (ms, R * M / ms * 1e3, ms), flush=True)
'''
for R, warps, ctas in [(24, 4, 1), (24, 16, 1), (24, 32, 1), (148, 4, 1), (148, 16, 1), (148, 32, 1),
                       (296, 16, 2), (888, 4, 6)]:
    env = dict(os.environ, PZ_CTA_WARPS=str(warps), PZ_G32_CTAS=str(ctas), PZ_PIPELINE="0")
    out = subprocess.run([sys.executable, "-c", child, L, str(R)], env=env, capture_output=True, text=True)
    print("R %4d  warps/CTA %2d  CTAs/SM %d: %s" % (R, warps, ctas, (out.stdout.strip().splitlines() or [out.stderr[-400:]])[-1]), flush=True)
