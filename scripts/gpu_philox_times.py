"""The bucketed Philox shuffle alone and under the fused step.  Usage: python scripts/gpu_philox_times.py [runs]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pypercolate_b200 import _native, lowering
runs = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
g = lowering.lowered_spanning_2d_grid(256); M = g.num_edges
ctx = _native.context_for(g, 0); ctx.set_ps(np.linspace(0.45, 0.55, 100))
seeds = (np.arange(runs, dtype=np.uint64) * 2654435761 % 2 ** 32).astype(np.uint32)
buf = torch.empty((runs, M), dtype=torch.int32, device="cuda")
mode = _native.RNG_MODES["philox"]
ctx.make_perms(2000, mode, seeds[:2000], out_device_ptr=buf.data_ptr()); ctx.synchronize()
t0 = time.perf_counter(); ctx.make_perms(runs, mode, seeds, out_device_ptr=buf.data_ptr()); ctx.synchronize()
dt = time.perf_counter() - t0
print("slab_kb=%s perms philox %8.1f ms  %.3e bonds/s" % (os.environ.get("PZ_PHILOX_SLAB_KB"), dt * 1e3, runs * M / dt), flush=True)
del buf; torch.cuda.empty_cache()
sd = torch.from_numpy(seeds.view(np.int32)).cuda()
for rep in range(2):
    ctx.reset_accumulators(); ctx.profile(True); ctx.timer_start()
    ctx.run_fused(runs, mode | _native.SEEDS_ON_DEVICE, sd.data_ptr(), _native.FUSE_MICRO | _native.FUSE_CANON)
    ms = ctx.timer_stop(); ph = ctx.profile_read(); ctx.profile(False)
print("fused philox PIPE=%s %8.1f ms  %.3e bonds/s  %s" % (os.environ.get("PZ_PIPELINE"), ms, runs * M / ms * 1e3, {k: round(v[0], 1) for k, v in ph.items() if v[1]}), flush=True)
