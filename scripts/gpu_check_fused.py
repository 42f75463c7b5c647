"""Scratch GPU check of the fused kernels vs the oracle."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pypercolate_b200 import _native, lowering
from oracle import oracle

bad = 0
def ok(cond, msg):
    global bad
    print(("OK   " if cond else "FAIL ") + msg)
    if not cond: bad += 1

# 1/2: device RNG
for L in (3, 8, 32, 128):
    g = lowering.lowered_spanning_2d_grid(L)
    ctx = _native.Context(0); ctx.set_graph(g)
    seeds = np.array([42, 0, 2**32 - 1, 3939566288, 7, 8, 9, 10, 11], dtype=np.uint32)
    pm = ctx.make_perms(seeds.size, _native.PERM_MT19937, seeds)
    ok(all(np.array_equal(pm[i], np.random.RandomState(int(s)).permutation(g.num_edges)) for i, s in enumerate(seeds)), "mt19937 L=%d" % L)
    pp = ctx.make_perms(seeds.size, _native.PERM_PHILOX, seeds)
    ok(all(np.array_equal(pp[i], oracle.philox_permutation(int(s), g.num_edges)) for i, s in enumerate(seeds)), "philox L=%d" % L)
    ctx.close()
g = lowering.lowered_spanning_2d_grid(256)
ctx = _native.Context(0); ctx.set_graph(g)
seeds = np.arange(5, dtype=np.uint32) + 100
pm = ctx.make_perms(5, _native.PERM_MT19937, seeds)
ok(all(np.array_equal(pm[i], oracle.numpy_permutation(int(s), g.num_edges)) for i, s in enumerate(seeds)), "mt19937 L=256")
pp = ctx.make_perms(5, _native.PERM_PHILOX, seeds)
ok(all(np.array_equal(pp[i], oracle.philox_permutation(int(s), g.num_edges)) for i, s in enumerate(seeds)), "philox L=256")
ctx.close()

# 3: fused micro accumulators
def limbs_value(words):
    return int(words[0]) + (int(words[1]) << 32)
for (g, R) in ((lowering.lowered_spanning_2d_grid(8), 70), (lowering.lowered_spanning_2d_grid(32), 40),
               (lowering.lowered_spanning_3d_grid(6), 33)):
  for force in (None, 2):
    if force is None: os.environ.pop("PZ_FORCE_STORE", None)
    else: os.environ["PZ_FORCE_STORE"] = str(force)
    ctx = _native.Context(0); ctx.set_graph(g)
    N, M = g.num_nodes, g.num_edges
    perms = np.stack([oracle.numpy_permutation(500 + r, M) for r in range(R)])
    rows = [oracle.sweep_rows(N, M, g.eu, g.ev, g.side_mask, g.preconnected, perms[r]) for r in range(R)]
    ctx.run_fused(R, _native.PERM_HOST, perms, _native.FUSE_MICRO)
    acc = ctx.micro_export()
    mx = np.stack([r['max_cluster_size'] for r in rows]).astype(object)
    mom = np.stack([r['moments'] for r in rows]).astype(object)   # R, M+1, 5
    span = np.stack([r['has_spanning_cluster'] for r in rows]).astype(np.int64)
    good = True
    good &= np.array_equal(np.cumsum(acc[:, 0].astype(np.int64)), span.sum(axis=0))
    good &= all(int(acc[n, 1]) == sum(mx[:, n]) for n in range(M + 1))
    good &= all(limbs_value(acc[n, 2:4]) == sum(x * x for x in mx[:, n]) for n in range(M + 1))
    cc = (N - 1) - mom[:, :, 0]
    good &= all(int(acc[n, 4]) == sum(cc[:, n]) for n in range(M + 1))
    good &= all(limbs_value(acc[n, 5:7]) == sum(x * x for x in cc[:, n]) for n in range(M + 1))
    for k in range(3):
        q = acc[:, 7 + 6 * k: 13 + 6 * k]
        for n in range(M + 1):
            s1 = int(q[n, 0]) + (int(q[n, 1]) << 32)
            s2 = int(q[n, 2]) + (int(q[n, 3]) << 32) + (int(q[n, 4]) << 64) + (int(q[n, 5]) << 96)
            good &= s1 == sum(mom[:, n, 2 + k]) and s2 == sum(x * x for x in mom[:, n, 2 + k])
    ok(good, "micro accumulators N=%d R=%d force=%s" % (N, R, force))
    mean, var = ctx.micro_finalize()
    fm = np.stack([r['max_cluster_size'] for r in rows]).astype(np.float64)
    fmom = np.stack([r['moments'] for r in rows]).astype(np.float64)
    good = np.array_equal(mean[0], span.sum(axis=0))
    good &= np.allclose(mean[1], fm.mean(axis=0), rtol=1e-13, atol=0)
    good &= np.allclose(var[0], fm.var(axis=0, ddof=1), rtol=1e-11, atol=0)
    for k in range(5):
        good &= np.allclose(mean[2 + k], fmom[:, :, k].mean(axis=0), rtol=1e-13, atol=0)
        v = fmom[:, :, k].var(axis=0, ddof=1)
        good &= np.allclose(var[1 + k], v, rtol=1e-9, atol=0)
        good &= np.array_equal(var[1 + k] == 0, fmom[:, :, k].std(axis=0, ddof=1) == 0)
    ok(good, "micro finalize N=%d R=%d force=%s" % (N, R, force))

    # 4-6: canonical
    ps = np.concatenate([np.linspace(0.0, 1.0, 11), np.linspace(0.45, 0.55, 7)])
    pmf = ctx.set_ps(ps, want_pmf=True)
    ref = np.stack([oracle.binomial_pmf(M, p) for p in ps])
    err = np.abs(pmf - ref).max() / ref.max()
    ok(err < 1e-14, "pmf M=%d maxerr %.2e" % (M, err))
    cols = np.stack([fm.mean(axis=0), fmom[:, :, 2].mean(axis=0), fmom[:, :, 4].mean(axis=0)])
    conv = ctx.convolve(cols)
    refc = cols @ ref.T
    ok(np.allclose(conv, refc, rtol=1e-12, atol=0), "convolve relerr %.2e" % np.abs(conv / refc - 1).max())
    ctx.reset_accumulators()
    ctx.run_fused(R, _native.PERM_HOST, perms, _native.FUSE_CANON | _native.FUSE_MICRO)
    per = ctx.canon_last_runs(R)
    refper = np.zeros_like(per)
    for r in range(R):
        for i, p in enumerate(ps):
            st = oracle.bond_canonical_statistics(rows[r], ref[i])
            refper[r, i, 0] = st['percolation_probability'][0]
            refper[r, i, 1] = st['max_cluster_size'][0]
            refper[r, i, 2:] = st['moments'][0]
    denom = np.maximum(np.abs(refper), 1e-300)
    rel = np.abs(per - refper) / denom
    rel[refper == 0] = np.abs(per[refper == 0])
    ok(rel.max() < 1e-11, "canon per-run maxrel %.2e" % rel.max())
    cnt, cm, cm2 = ctx.canon_export()
    ok(cnt == R and np.allclose(cm, refper.mean(axis=0), rtol=1e-12, atol=1e-300)
       and np.allclose(cm2, ((refper - refper.mean(axis=0)) ** 2).sum(axis=0), rtol=1e-8, atol=1e-12 * np.abs(refper).max() ** 2),
       "canon reduce")
    st1 = ctx.canonical_statistics_rows(rows[0], ref[5])
    ok(np.allclose(st1, refper[0, 5], rtol=1e-12, atol=1e-300), "canonical_statistics_rows")
    ctx.close()
os.environ.pop("PZ_FORCE_STORE", None)

# 7: timing at L=256 / L=128 (wall clock around synchronised calls)
for L, R in ((128, 4096), (256, 2048)):
    g = lowering.lowered_spanning_2d_grid(L)
    ctx = _native.Context(0); ctx.set_graph(g)
    M = g.num_edges
    seeds = np.arange(R, dtype=np.uint32)
    import torch
    dperm = torch.empty((R, M), dtype=torch.int32, device="cuda")
    for mode, name in ((_native.PERM_PHILOX, "philox"), (_native.PERM_MT19937, "mt19937")):
        ctx.make_perms(R, mode, seeds, out_device_ptr=dperm.data_ptr()); ctx.synchronize()
        t0 = time.time(); ctx.make_perms(R, mode, seeds, out_device_ptr=dperm.data_ptr()); ctx.synchronize()
        dt = time.time() - t0
        print("L=%d R=%d perms %s: %.1f ms  (%.3g bonds/s)" % (L, R, name, dt * 1e3, R * M / dt))
    ctx.set_ps(np.linspace(0.45, 0.55, 100))
    for flags, name in ((_native.FUSE_MICRO, "micro"), (_native.FUSE_MICRO | _native.FUSE_CANON, "micro+canon")):
        ctx.reset_accumulators()
        ctx.run_fused(R, _native.PERM_DEVICE, dperm.data_ptr(), flags); ctx.synchronize()
        t0 = time.time(); ctx.run_fused(R, _native.PERM_DEVICE, dperm.data_ptr(), flags); ctx.synchronize()
        dt = time.time() - t0
        print("L=%d R=%d fused %s (perms resident): %.1f ms  (%.3g bonds/s)" % (L, R, name, dt * 1e3, R * M / dt))
    ctx.close()
print("TOTAL BAD", bad)
sys.exit(1 if bad else 0)
