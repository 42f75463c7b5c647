"""Whole fused step per bond-order generator (argv) under PZ_PIPELINE / PZ_FINDERS / ... from the environment; RUNS=n."""
import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from pypercolate_b200 import _native, lowering
runs = int(os.environ.get('RUNS', '20000'))
g = lowering.lowered_spanning_2d_grid(256); M = g.num_edges
ctx = _native.context_for(g, 0); ctx.set_ps(np.linspace(0.45, 0.55, 100))
seeds = (np.arange(runs, dtype=np.uint64) * 2654435761 % 2 ** 32).astype(np.uint32)
sd = torch.from_numpy(seeds.view(np.int32)).cuda()
for name in sys.argv[1:]:
    mode = _native.RNG_MODES[name] | _native.SEEDS_ON_DEVICE
    for rep in range(3):
        ctx.reset_accumulators(); ctx.profile(True); ctx.timer_start()
        ctx.run_fused(runs, mode, sd.data_ptr(), _native.FUSE_MICRO | _native.FUSE_CANON)
        ms = ctx.timer_stop(); ph = ctx.profile_read(); ctx.profile(False)
    print("PIPE=%s %-10s %8.1f ms  %.3e bonds/s  %s" % (os.environ.get("PZ_PIPELINE"), name, ms, runs * M / ms * 1e3, {k: round(v[0], 1) for k, v in ph.items() if v[1]}), flush=True)
