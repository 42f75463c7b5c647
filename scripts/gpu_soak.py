"""Soak: many full L = 256 runs through the shipped sweep (finder warps) and through the plain lock-step kernel
(PZ_FINDERS=0) on the same bond orders -- the exact per-n sums and the canonical partials must be identical --
for several seed sets and every device generator.  Usage: python scripts/gpu_soak.py [runs per set] [sets]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pypercolate_b200 import _native, lowering

runs = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
sets = int(sys.argv[2]) if len(sys.argv) > 2 else 4
g = lowering.lowered_spanning_2d_grid(256)
ps = np.linspace(0.45, 0.55, 9)
flags = _native.FUSE_MICRO | _native.FUSE_CANON
bad = 0
t0 = time.time()
for k in range(sets):
    rng = ["philox", "mt19937", "feistel", "philox_fy"][k % 4]
    seeds = ((np.arange(runs, dtype=np.uint64) + 1000003 * (k + 1)) * 2654435761 % 2 ** 32).astype(np.uint32)
    got = {}
    for finders in ("1", "0"):
        os.environ["PZ_FINDERS"] = finders
        ctx = _native.Context(0); ctx.set_graph(g); ctx.set_ps(ps)
        ctx.run_fused(runs, _native.RNG_MODES[rng], seeds, flags)
        got[finders] = (ctx.micro_export(), ctx.canon_export())
        ctx.close()
    same = (np.array_equal(got["1"][0], got["0"][0]) and got["1"][1][0] == got["0"][1][0]
            and np.array_equal(got["1"][1][1], got["0"][1][1]) and np.array_equal(got["1"][1][2], got["0"][1][2]))
    print("set %d (%s, %d runs): %s" % (k, rng, runs, "identical" if same else "MISMATCH"), flush=True)
    bad += not same
os.environ.pop("PZ_FINDERS", None)
print("%d sets, %d mismatches, %.1f s" % (sets, bad, time.time() - t0))
sys.exit(1 if bad else 0)
