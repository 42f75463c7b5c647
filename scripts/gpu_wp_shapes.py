"""Warp-per-run bond-order kernels: CTAs per SM x warps per CTA, alone and under the fused step.
Usage: python scripts/gpu_wp_shapes.py [runs] [rng]   (each shape runs in a fresh process: the knobs are read once)"""
import os, subprocess, sys
runs = sys.argv[1] if len(sys.argv) > 1 else "8000"
rng = sys.argv[2] if len(sys.argv) > 2 else "mt19937"
child = r'''
import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from pypercolate_b200 import _native, lowering
runs, rng = int(sys.argv[1]), sys.argv[2]
g = lowering.lowered_spanning_2d_grid(256); M = g.num_edges
ctx = _native.context_for(g, 0); ctx.set_ps(np.linspace(0.45, 0.55, 100))
seeds = (np.arange(runs, dtype=np.uint64) * 2654435761 % 2 ** 32).astype(np.uint32)
buf = torch.empty((runs, M), dtype=torch.int32, device="cuda")
mode = _native.RNG_MODES[rng]
ctx.make_perms(600, mode, seeds[:600], out_device_ptr=buf.data_ptr()); ctx.synchronize()
t0 = time.perf_counter(); ctx.make_perms(runs, mode, seeds, out_device_ptr=buf.data_ptr()); ctx.synchronize()
dt = time.perf_counter() - t0
del buf; torch.cuda.empty_cache()
sd = torch.from_numpy(seeds.view(np.int32)).cuda()
for rep in range(2):
    ctx.reset_accumulators(); ctx.profile(True); ctx.timer_start()
    ctx.run_fused(runs, mode | _native.SEEDS_ON_DEVICE, sd.data_ptr(), _native.FUSE_MICRO | _native.FUSE_CANON)
    ms = ctx.timer_stop(); ph = ctx.profile_read(); ctx.profile(False)
print("alone %7.1f ms %.3e bonds/s | fused %7.1f ms %.3e bonds/s %s" % (dt * 1e3, runs * M / dt, ms, runs * M / ms * 1e3,
      {k: round(v[0], 1) for k, v in ph.items() if v[1] and k in ("perm", "sweep")}), flush=True)
'''
for ctas, warps in [(1, 1), (1, 2), (1, 4), (1, 8), (2, 8)]:
    env = dict(os.environ, PZ_WP_CTAS=str(ctas), PZ_WP_WARPS=str(warps))
    out = subprocess.run([sys.executable, "-c", child, runs, rng], env=env, capture_output=True, text=True)
    print("ctas/SM %d warps %d: %s" % (ctas, warps, (out.stdout.strip().splitlines() or [out.stderr[-300:]])[-1]), flush=True)
