"""Small target for ncu captures: [what] = philox | mt19937 | philox_fy | fused (default) [runs] [L]."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pypercolate_b200 import _native, lowering
what = sys.argv[1] if len(sys.argv) > 1 else "fused"
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 592
L = int(sys.argv[3]) if len(sys.argv) > 3 else 256
g = lowering.lowered_spanning_2d_grid(L); M = g.num_edges
ctx = _native.context_for(g, 0); ctx.set_ps(np.linspace(0.45, 0.55, 100))
seeds = (np.arange(runs, dtype=np.uint64) * 2654435761 % 2 ** 32).astype(np.uint32)
if what == "fused":
    rng = os.environ.get("PZ_RNG", "philox")
    for rep in range(2):
        ctx.reset_accumulators()
        ctx.run_fused(runs, _native.RNG_MODES[rng], seeds, _native.FUSE_MICRO | _native.FUSE_CANON)
else:
    buf = torch.empty((runs, M), dtype=torch.int32, device="cuda")
    for rep in range(2):
        ctx.make_perms(runs, _native.RNG_MODES[what], seeds, out_device_ptr=buf.data_ptr())
ctx.synchronize()
