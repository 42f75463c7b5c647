#!/usr/bin/env python
"""
Headline benchmark: bond-additions/s of the fused Newman-Ziff hot path on the
L = 256 square lattice (BASELINE.json config 3: 1e5 runs, fused microcanonical
+ canonical averaging at 100 p values, runs sharded over the GPUs of one box).

    python bench.py --gpus N --steps K --warmup W            # our arm
    python bench.py --impl reference --gpus N --steps K ...  # CPU reference arm

A step is one pass of the hot path over the whole batch of runs: bond orders
(device RNG, `--rng`: numpy's MT19937 stream bit for bit, Philox Fisher-Yates, the
bucketed Philox shuffle or the Philox-keyed Feistel bijection) -> union-find sweep ->
per-n exact sums over runs and per-run binomial contraction -> (multi-GPU) one NCCL
exchange -> per-n mean / variance.  Strong scaling: the 1e5 runs are split over the ranks.

One JSON line is printed by rank 0 (see the driver contract in the task text).  Besides the
headline (BASELINE config 3) it carries, under "configs", one short measurement of the other
GPU configs of BASELINE.json (c2, c4, c5, sharded over the same N GPUs) and of config 3 under
the other generators, under "multi_gpu_parity" (N > 1) the comparison of a sharded job with the
same job on one rank, and under "e2e" the cold first call next to the warm figure.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

L = 256
TOTAL_RUNS = int(os.environ.get("PZ_BENCH_RUNS", "100000"))
NUM_P = 100
METRIC = "bond-additions/sec (L=256 square lattice, fused microcanonical + canonical averaging)"
UNIT = "bond-additions/s"
# algorithmic bytes per bond addition of the fused 2D path:
# 4 (bond order) + 8 (endpoints) + 8 (two root look-ups) + 8 * (N-1)/M (merge writes)
ALGO_BYTES_PER_BOND = 20.0 + 8.0 * (L * L - 1) / (2 * L * (L - 1))


RNG_TEXT = {
    "feistel": "bond orders on device = Philox-keyed 20-round Feistel bijection with cycle walking",
    "philox": "bond orders on device = counter-based Philox4x32-10 Fisher-Yates in shared memory (buckets of ~64 bonds, exact uniform shuffle)",
    "mt19937": "bond orders on device = numpy RandomState(seed).permutation stream, bit for bit",
    "philox_fy": "bond orders on device = textbook Fisher-Yates with Philox4x32-10 counter-based draws",
}


# the north star's device generator: counter-based Philox, an exact uniform shuffle (bucketed
# Fisher-Yates in shared memory).  --rng mt19937 is the mode in which every run is bit-comparable
# with the reference (numpy's stream); every generator is measured under "configs".
DEFAULT_RNG = "philox"


def workload_config(n_gpus, runs_total, rng=DEFAULT_RNG):
    return {
        "workload": "spanning_2d_grid L=256 (N=65536, M=130560), %d runs total, %s, "
                    "fused microcanonical sums + per-run canonical contraction at "
                    "%d p in [0.45, 0.55]" % (runs_total, RNG_TEXT[rng], NUM_P),
        "rng": rng,
        "runs_total": runs_total,
        "runs_per_gpu": runs_total // n_gpus,
        "num_p": NUM_P,
        "parallelism": "runs sharded over %d GPU(s), one NCCL exchange per step" % n_gpus,
        "l2_policy": "inputs larger than L2 (bond orders + merge records of a launch: several GB)",
    }


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + q,
                                       "--format=csv,noheader,nounits", "-lms", "200"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
            except ValueError:
                continue
            for name, v in zip(names, parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        self.f.close()
        os.unlink(self.f.name)
        if sm:
            out["sm_mhz"] = float(np.median(sm))
            out["sm_max_mhz"] = float(max(smax))
        out["reasons"] = sorted(reasons)
        out["samples"] = len(sm)
        return out


def cpu_fused_rate(seconds_target):
    """The reference's algorithm on the host cores: the oracle port (oracle/pz_oracle.c) of the
    WHOLE fused job of config 3 -- numpy-stream permutation, sweep, per-n sums over runs, per-run
    binomial contraction at 100 p folded into (count, mean, M2) -- OpenMP over every host thread.
    The thread count is set explicitly (a multi-process launcher pre-sets OMP_NUM_THREADS=1) and
    the count OpenMP really used is what is reported."""
    from oracle import oracle
    from pypercolate_b200 import lowering
    g = lowering.lowered_spanning_2d_grid(L)
    want = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    threads = oracle.set_threads(want)
    global _CPU_TABLES
    if _CPU_TABLES is None:      # binomial weights: built once, like pz_set_ps keeps them on the GPU
        _CPU_TABLES = oracle.fused_tables(g.num_edges, np.linspace(0.45, 0.55, NUM_P))
    ps = _CPU_TABLES
    warm = np.arange(threads, dtype=np.uint32) + 1
    t0 = time.perf_counter()
    oracle.fused_many(g.num_nodes, g.num_edges, g.eu, g.ev, g.side_mask, False, warm, ps)
    per_wave = max(time.perf_counter() - t0, 1e-3)
    waves = max(1, int(seconds_target / per_wave))
    seeds = np.arange(threads * waves, dtype=np.uint32) + 1000
    t0 = time.perf_counter()
    micro, mean, m2 = oracle.fused_many(g.num_nodes, g.num_edges, g.eu, g.ev, g.side_mask, False, seeds, ps)
    dt = time.perf_counter() - t0
    assert micro[1, -1] == float(seeds.size) * g.num_nodes        # every run ends in one cluster
    return seeds.size * g.num_edges / dt, threads, seeds.size, dt


_CPU_TABLES = None
CPU_SAMPLE_TEXT = ("%d runs of L=256 in %.1f s: numpy-stream permutation + sweep + per-n sums over runs + "
                   "per-run binomial contraction at 100 p folded into (count, mean, M2); C port of the "
                   "reference's algorithm (oracle/pz_oracle.c), OpenMP over %d threads (measured: "
                   "omp_get_num_threads inside a parallel region)")


def run_reference(args, rank, world):
    if rank != 0:
        return
    vals = []
    per_step = max(2.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
    cores = runs = 0
    for i in range(args.warmup + args.steps):
        rate, cores, runs, dt = cpu_fused_rate(per_step)
        if i >= args.warmup:
            vals.append((rate, dt))
    value = float(np.mean([v for v, _ in vals]))
    dt = float(np.mean([d for _, d in vals]))
    config = workload_config(args.gpus, TOTAL_RUNS, "mt19937")
    config["timed_sample"] = ("each step times %d runs of this workload (same lattice, same generator, "
                              "same statistics) on %d host threads; bond-additions/s does not depend on "
                              "the number of runs (independent units)" % (runs, cores))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": config,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": CPU_SAMPLE_TEXT % (runs, dt, cores)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the reference is pure Python (~4e4 bond-additions/s per core, BASELINE.md) and "
                "cannot travel to the GPU box; this arm times the C restatement of its algorithm "
                "(oracle/pz_oracle.c) doing the same job as the GPU arm, on all host threads",
    }
    emit(line)


def emit(line):
    """The ONE JSON line of the contract, on the real stdout."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


# everything else a library may print (NCCL's version banner goes to fd 1) lands on stderr
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


SECONDARY = [
    # name, lattice, L, total runs, (p_lo, p_hi, num_p), flags, rng, what BASELINE.json says
    ("c2", "2d", 128, 10 ** 4, (0.4, 0.6, 40), "canon", None,
     "spanning_2d_grid L=128, 1e4 runs, per-run statistics reduced with bond_reduce at the jugfile's 40 p "
     "(parent array in shared memory)"),
    ("c4", "2d", 1024, 10 ** 3, (0.45, 0.55, 100), "both", None,
     "spanning_2d_grid L=1024, 1e3 runs, fused microcanonical + canonical averaging (global-memory store)"),
    ("c5", "3d", 64, 10 ** 4, (0.2, 0.3, 100), "both", None,
     "simple-cubic L=64 as a general edge list with top/bottom sides, 1e4 runs, fused microcanonical + "
     "canonical averaging (global-memory store)"),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--runs", type=int, default=TOTAL_RUNS, help="total runs per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true",
                    help="skip the short measurements of the other configs / generators")
    ap.add_argument("--rng", default=DEFAULT_RNG, choices=sorted(RNG_TEXT),
                    help="device bond-order generator of our arm")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import __graft_entry__
    __graft_entry__.build()
    from pypercolate_b200 import _native, hpc, lowering, multi

    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world,
                                device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    def seeds_of(lo, hi):
        return (np.arange(lo, hi, dtype=np.uint64) * 2654435761 % (2 ** 32)).astype(np.uint32)

    g = lowering.lowered_spanning_2d_grid(L)
    M, N = g.num_edges, g.num_nodes
    ctx = _native.context_for(g, local)           # CUDA context + graph upload: not part of any timing
    ps = np.linspace(0.45, 0.55, NUM_P)
    lo, hi = multi.shard_bounds(args.runs, rank, world)
    my_runs = hi - lo
    seeds_host = seeds_of(lo, hi)
    flags = _native.FUSE_MICRO | _native.FUSE_CANON
    mode = _native.RNG_MODES[args.rng] | _native.SEEDS_ON_DEVICE

    def step_e2e():
        """the public call: host seeds / ps in, host statistics out"""
        return hpc.bond_statistics_batch(g, N, M, seeds_host, ps, 0.3173, device=local,
                                         rng=args.rng, distributed=world > 1)

    # ---- the first call of a study: nothing cached (binomial table, beta-interval table, scratch
    # buffers, NCCL communicator) ---------------------------------------------------------------
    barrier()
    t0 = time.perf_counter()
    step_e2e()
    barrier()
    e2e_cold_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)

    seeds_dev = torch.from_numpy(seeds_host.view(np.int32)).cuda()
    ctx.set_ps(ps)

    def step_device():
        """inputs (graph, seeds, binomial weights) resident in HBM"""
        ctx.reset_accumulators()
        ctx.run_fused(my_runs, mode, seeds_dev.data_ptr(), flags)
        if world > 1:
            multi.allreduce_context(ctx)
        ctx.micro_finalize_device_only()

    # ---- device-resident throughput -------------------------------------------
    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    ctx.profile(True)
    launches0 = ctx.launch_count
    ctx.timer_start()
    for _ in range(args.steps):
        step_device()
    ms = ctx.timer_stop()
    barrier()
    launches = ctx.launch_count - launches0
    phases = ctx.profile_read()
    ctx.profile(False)
    clocks = sampler.stop() if sampler else None

    # ---- end to end through the public API (warm: tables of the first call are kept) ----------
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0

    t = torch.tensor([ms, e2e_s * 1e3, float(launches)], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, e2e_ms, launches = float(tmax[0]), float(tmax[1]), int(tsum[2])
    else:
        e2e_ms = e2e_s * 1e3
    assert res["number_of_runs"] == args.runs

    # ---- multi-GPU parity (untimed): a sharded job against the same job on rank 0 alone ---------
    parity = None
    if world > 1:
        pr = 1200
        plo, phi = multi.shard_bounds(pr, rank, world)
        ctx.reset_accumulators()
        ctx.run_fused(phi - plo, _native.RNG_MODES[args.rng], seeds_of(plo, phi), flags)
        multi.allreduce_context(ctx)
        sharded_runs = ctx.micro_runs
        sh_mean, sh_var = ctx.micro_finalize()
        sh_canon = ctx.canon_export()
        if rank == 0:
            ctx.reset_accumulators()
            ctx.run_fused(pr, _native.RNG_MODES[args.rng], seeds_of(0, pr), flags)
            one_mean, one_var = ctx.micro_finalize()
            one_canon = ctx.canon_export()
            rel = lambda a, b: float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))
            # M2 is a sum of squared deviations: where the runs agree to many digits (the spanning
            # probability at a p where every run is 1 - 1e-12) it is rounding noise on both sides, so
            # its error is judged against max(M2, 1e-6 count mean^2) -- a wrong merge moves a
            # well-conditioned M2 by O(M2) and is caught either way
            scale = np.abs(one_canon[2]) + 1e-6 * pr * one_canon[1] ** 2
            parity = {
                "runs": pr, "ranks": world,
                "micro_sums_bit_equal": bool(sharded_runs == pr and np.array_equal(sh_mean, one_mean)
                                             and np.array_equal(sh_var, one_var)),
                "canonical_count_equal": bool(sh_canon[0] == one_canon[0] == pr),
                "canonical_mean_max_rel": rel(sh_canon[1], one_canon[1]),
                "canonical_m2_max_rel": rel(sh_canon[2], one_canon[2]),
                "canonical_m2_max_err_over_max_of_m2_and_1e-6_count_mean2":
                    float(np.max(np.abs(sh_canon[2] - one_canon[2]) / np.maximum(scale, 1e-300))),
            }
            parity["ok"] = bool(parity["micro_sums_bit_equal"] and parity["canonical_count_equal"]
                                and parity["canonical_mean_max_rel"] < 1e-13
                                and parity["canonical_m2_max_err_over_max_of_m2_and_1e-6_count_mean2"] < 1e-10)
        barrier()

    # ---- the other configs and generators: one short measurement each --------------------------
    # (warm-up pass, then one timed pass; same sharding, exchange and finalize as the headline)
    def close_context(graph):
        c = graph._handles.pop(local, None)
        if c is not None:
            c.close()

    def measure(graph, runs_total, rng, p_range, what):
        c = _native.context_for(graph, local)
        fl = {"canon": _native.FUSE_CANON, "both": _native.FUSE_MICRO | _native.FUSE_CANON}[what]
        c.set_ps(np.linspace(*p_range))
        a, b = multi.shard_bounds(runs_total, rank, world)
        sd = torch.from_numpy(seeds_of(a, b).view(np.int32)).cuda()
        md = _native.RNG_MODES[rng] | _native.SEEDS_ON_DEVICE
        out = None
        for timed in (False, True):
            barrier()
            c.profile(timed)
            c.timer_start()
            c.reset_accumulators()
            if b > a:
                c.run_fused(b - a, md, sd.data_ptr(), fl)
            if world > 1:
                multi.allreduce_context(c)
            if fl & _native.FUSE_MICRO:
                c.micro_finalize_device_only()
            t_ms = c.timer_stop()
            barrier()
            out = max_over_ranks(t_ms)
        ph = c.profile_read()
        c.profile(False)
        bonds = float(runs_total) * graph.num_edges
        return {"ms": out, "bonds_per_s": bonds / (out * 1e-3), "runs_per_s": runs_total / (out * 1e-3),
                "runs": runs_total, "rng": rng, "n_gpus": world,
                "sweep_bonds_per_s_rank0": (float(b - a) * graph.num_edges / (ph["sweep"][0] * 1e-3)
                                            if ph["sweep"][1] else None)}

    configs = {}
    if not args.no_secondary:
        for other in ("mt19937", "philox_fy", "philox", "feistel"):
            if other == args.rng:
                continue
            r = measure(g, args.runs, other, (0.45, 0.55, NUM_P), "both")
            r["workload"] = "config 3 (the headline workload) with " + RNG_TEXT[other]
            configs["c3_" + other] = r
        close_context(g)
        for name, kind, size, runs_total, p_range, what, rng, text in SECONDARY:
            graph = (lowering.lowered_spanning_2d_grid if kind == "2d" else lowering.lowered_spanning_3d_grid)(size)
            r = measure(graph, runs_total, rng or args.rng, p_range, what)
            r["workload"] = text
            configs[name] = r
            close_context(graph)

    if rank == 0:
        bonds_per_step = float(args.runs) * M
        value = bonds_per_step * args.steps / (ms * 1e-3)
        e2e_value = bonds_per_step * args.steps / (e2e_ms * 1e-3)
        sweep_ms, sweep_launches = phases["sweep"]
        peak, peak_src = measured_peak()
        bonds_per_launch = float(my_runs) * M * args.steps / max(1, sweep_launches)
        achieved = ALGO_BYTES_PER_BOND * bonds_per_launch / (sweep_ms / max(1, sweep_launches) * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "sweep_traffic.json")
        if os.path.exists(tpath):
            try:
                tj = json.load(open(tpath))
                traffic = float(tj["dram_bytes_per_bond"]) * bonds_per_launch
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": workload_config(world, args.runs, args.rng),
            "runs_per_s": float(args.runs) * args.steps / (ms * 1e-3),
            "roofline": {
                "bound": "hbm", "kernel": "sweep_fw_kernel (union-find sweep, 16 main + 4 finder warps per run; shared-memory latency bound, see DESIGN.md)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src,
                "traffic_source": "profiles/sweep_traffic.json: DRAM bytes read + written by one ncu-captured launch "
                                  "(592 runs), per bond, x the bonds of one launch here",
                "algorithmic_bytes_per_bond": ALGO_BYTES_PER_BOND,
                "avg_launch_ms": sweep_ms / max(1, sweep_launches),
                "bonds_per_launch": bonds_per_launch,
            },
            "phase_ms_per_step_rank0": {k: v[0] / args.steps for k, v in phases.items() if v[1]},
            "e2e": {
                "value": e2e_value, "unit": UNIT,
                "h2d_bytes_per_step": int(seeds_host.nbytes + ps.nbytes),
                "d2h_bytes_per_step": int(19 * (M + 1) * 8 + 2 * NUM_P * 7 * 8),   # pz_micro_arrays + canonical partials
                "ms_per_step": e2e_ms / args.steps,
                "cold_first_call_ms": e2e_cold_ms,
                "cold_note": "first call in the process: binomial table (100 p), beta-interval table, scratch "
                             "allocation and (N > 1) the NCCL communicator are built inside it; the warm figure "
                             "re-uses them, as every later call of a study does",
                "api": "pypercolate_b200.hpc.bond_statistics_batch (host seeds/ps in, "
                       "microcanonical arrays + finalized canonical averages out)",
            },
            "gpu_launches": launches,
            "clocks": clocks,
            "configs": configs,
        }
        if parity is not None:
            line["multi_gpu_parity"] = parity
        if not args.no_cpu_baseline and world == 1:
            rate, cores, nruns, dt = cpu_fused_rate(12.0)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": CPU_SAMPLE_TEXT % (nruns, dt, cores)}
        emit(line)
        if parity is not None and not parity["ok"]:
            raise SystemExit("multi-GPU parity check failed: %r" % (parity,))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
