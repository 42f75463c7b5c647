import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from pypercolate_b200 import _native, lowering
for kind, L, runs, pr in (("2d", 1024, 1000, (0.45, 0.55, 100)), ("3d", 64, 10000, (0.2, 0.3, 100))):
    g = (lowering.lowered_spanning_2d_grid if kind == "2d" else lowering.lowered_spanning_3d_grid)(L)
    M = g.num_edges
    ctx = _native.context_for(g, 0); ctx.set_ps(np.linspace(*pr))
    seeds = (np.arange(runs, dtype=np.uint64) * 2654435761 % 2 ** 32).astype(np.uint32)
    sd = torch.from_numpy(seeds.view(np.int32)).cuda()
    for name in sys.argv[1:]:
        mode = _native.RNG_MODES[name] | _native.SEEDS_ON_DEVICE
        for rep in range(2):
            ctx.reset_accumulators(); ctx.profile(True); ctx.timer_start()
            ctx.run_fused(runs, mode, sd.data_ptr(), _native.FUSE_MICRO | _native.FUSE_CANON)
            ms = ctx.timer_stop(); ph = ctx.profile_read(); ctx.profile(False)
        print("%s L=%d %-8s %8.1f ms  %.3e bonds/s  %s" % (kind, L, name, ms, runs * M / ms * 1e3, {k: round(v[0], 1) for k, v in ph.items() if v[1]}), flush=True)
    ctx.close()
