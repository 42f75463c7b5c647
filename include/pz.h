/*
 * pz.h -- C-ABI of the B200-native Newman-Ziff bond-percolation hot path.
 *
 * The reference (andsor/pypercolate) is pure Python and has NO plugin / FFI
 * layer: its boundary is the Python function surface of percolate/hpc.py and
 * percolate/percolate.py.  This header is the boundary a maintainer would
 * bind (ctypes stub in INTEGRATION.md); every entry point names the reference
 * function whose work it replaces (paths relative to the reference checkout).
 *
 * Conventions: plain pointers and sizes, caller-owned buffers, no torch
 * types.  Every call returns 0 on success or a negative pz_status; the
 * message of the last failure on the calling thread is pz_last_error().
 * A context owns one CUDA device, one stream and its scratch; calls on one
 * context are not re-entrant.  "host" pointers may be pageable or pinned.
 */
#ifndef PZ_H
#define PZ_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pz_ctx pz_ctx;

enum pz_status {
    PZ_OK = 0,
    PZ_ERR_ARG = -1,       /* bad argument                                  */
    PZ_ERR_CUDA = -2,      /* CUDA runtime failure (message has the detail)  */
    PZ_ERR_STATE = -3,     /* call order (no graph set, nothing accumulated) */
    PZ_ERR_NOMEM = -4
};

/* how the bond order of a run is obtained (percolate/hpc.py:195,206) */
enum pz_perm_mode {
    PZ_PERM_HOST = 0,      /* caller supplies int32 perms[R][M] (exact mode)             */
    PZ_PERM_DEVICE = 1,    /* same layout, pointer is device memory                      */
    PZ_PERM_MT19937 = 2,   /* uint32 seeds[R]; numpy RandomState(seed).permutation(M)
                              reproduced on the device bit for bit                      */
    PZ_PERM_PHILOX = 3,    /* uint32 seeds[R]; counter-based Philox4x32-10 Fisher-Yates in
                              shared memory (bucketed, exact uniform shuffle; M <= 2^24)  */
    PZ_PERM_FEISTEL = 4,   /* uint32 seeds[R]; Philox-keyed 20-round Feistel bijection of
                              [0, M) with cycle walking: order[n] = pi_seed(n), no
                              scratch and no shared memory                              */
    PZ_PERM_PHILOX_FY = 5  /* uint32 seeds[R]; textbook Fisher-Yates (i = M-1..1, swap with
                              j in [0, i]) with unbiased Lemire draws from Philox4x32-10
                              counters: one warp per run, no shared memory; like
                              PZ_PERM_MT19937 it is generated underneath the sweep of the
                              previous batch of runs                                    */
};

/* OR into a device-RNG perm_mode when the seeds already live in device memory */
#define PZ_SEEDS_ON_DEVICE 0x100

/* number of 64-bit words per bond count n in the micro accumulators */
#define PZ_ACC_WORDS 25
/* columns of a per-run canonical statistics row: P_span, max, moments[5] */
#define PZ_CANON_COLS 7

const char *pz_last_error(void);
int pz_version(void);

/* ---- context ------------------------------------------------------------ */
int pz_create(int device, pz_ctx **out);
void pz_destroy(pz_ctx *ctx);
int pz_device(const pz_ctx *ctx);
/* the CUDA stream (cudaStream_t as void*) all work of this context runs on */
void *pz_stream(const pz_ctx *ctx);
int pz_synchronize(pz_ctx *ctx);

/*
 * Upload the lowered graph: replaces the per-bond walks over
 * ``perc_graph.edges()`` / ``nodes_iter()`` and the auxiliary-node set-up of
 * bond_sample_states (percolate/hpc.py:205, 224-246).
 *   eu, ev      host int32[M], endpoints of bond e in perc_graph.edges() order
 *   side_mask   host uint8[N] (bit s: node touches spanning side s) or NULL
 *               for spanning_cluster=False
 *   preconnected  auxiliary structure alone joins both sides
 */
int pz_set_graph(pz_ctx *ctx, int32_t N, int32_t M, const int32_t *eu,
                 const int32_t *ev, const uint8_t *side_mask, int preconnected);

/* bytes of one packed microcanonical_statistics_dtype row
 * (percolate/hpc.py:57-70): 53 with spanning, 52 without */
int pz_row_bytes(const pz_ctx *ctx);

/*
 * R complete runs, materialised: replaces bond_sample_states /
 * bond_microcanonical_statistics (percolate/hpc.py:73-307, 310-404) and feeds
 * sample_states / single_run_arrays (percolate/percolate.py:103-447).
 *   perm_mode, perm_src   see pz_perm_mode
 *   rows_out   host, R * (M+1) * pz_row_bytes() bytes, packed rows; row 0 of
 *              each run has edge = 0 (undefined in the reference)
 *   perms_out  optional host int32[R][M]: the bond orders used (device RNG modes)
 */
int pz_run_rows(pz_ctx *ctx, int32_t R, int perm_mode, const void *perm_src,
                void *rows_out, int32_t *perms_out);

/*
 * Bond orders only: the device replacement of RandomState(seed).permutation(M)
 * (percolate/hpc.py:195,206).  perm_mode is any of the device RNG modes (PZ_PERM_MT19937 ..);
 * seeds host uint32[R]; out int32[R][M] (host, or device when is_device != 0).
 */
int pz_make_perms(pz_ctx *ctx, int32_t R, int perm_mode, const uint32_t *seeds,
                  int32_t *out, int is_device);

/*
 * Fused path: R runs are swept and folded, on the device, into
 *  (a) exact per-n integer sums over runs (micro accumulators) -- the inputs of
 *      _microcanonical_average_* (percolate/percolate.py:450-705), and/or
 *  (b) per-run canonical statistics sum_n f_p[n] Q[n]
 *      (bond_canonical_statistics, percolate/hpc.py:443-515) reduced over runs
 *      to (count, mean, M2) (bond_initialize_canonical_averages + bond_reduce,
 *      percolate/hpc.py:561-702).
 * Nothing per-run leaves the device.  Accumulators persist in the context
 * until pz_reset_accumulators(); successive calls add runs.
 *   flags  bit 0: accumulate (a);  bit 1: accumulate (b) (needs pz_set_ps first)
 */
#define PZ_FUSE_MICRO 1
#define PZ_FUSE_CANON 2
int pz_run_fused(pz_ctx *ctx, int32_t R, int perm_mode, const void *perm_src,
                 int flags);
int pz_reset_accumulators(pz_ctx *ctx);

/*
 * Micro accumulators: for n = 0..M, PZ_ACC_WORDS uint64 words.  Every sum is
 * held as 32-bit limbs in 64-bit words, so a WORD-WISE integer sum of the
 * blocks of several contexts (the cross-GPU all-reduce) is exact:
 *   [0]      runs whose spanning cluster FIRST appears at n (delta form;
 *            the count of spanning runs at n is the prefix sum)
 *   [1]      sum max            [2],[3]  sum max^2      (lo32, hi32)
 *   [4]      sum c              [5],[6]  sum c^2        (c = merges so far;
 *                                                        moments[0] = N-1-c)
 *   [7+6j ..12+6j], j = 0,1,2 for moments[2+j] = m:
 *            sum m (lo32, hi32), sum m^2 (four 32-bit limbs)
 * (moments[1] = N - max needs no words of its own.)
 * pz_micro_runs: number of runs folded so far.  Export copies the block to
 * caller memory (host, or device when is_device != 0); import replaces the
 * block and the run count (after an all-reduce).
 */
int64_t pz_micro_runs(const pz_ctx *ctx);
int pz_micro_export(pz_ctx *ctx, uint64_t *dst, int is_device);
int pz_micro_import(pz_ctx *ctx, const uint64_t *src, int is_device, int64_t runs);

/*
 * Per-n means and sample variances from the exact sums (float64), i.e. the
 * arithmetic of percolate/percolate.py:557-563, 613-620, 681-690 before the
 * scipy quantile calls:
 *   mean_out  host double[7][M+1]: spanning count k, max, moments[0..4]
 *   var_out   host double[6][M+1]: unbiased variance (ddof=1) of max, moments[0..4];
 *             exactly 0.0 when all runs agree
 * Both NULL: the arrays are computed and left on the device.
 */
int pz_micro_finalize(pz_ctx *ctx, double *mean_out, double *var_out);

/*
 * The arrays of microcanonical_averages_arrays (percolate/percolate.py:968-1064)
 * straight from the exact sums: per-n sample mean and Student-t interval
 * t * std / sqrt(runs) + mean, (mean, mean) when the sample std is zero
 * (percolate/percolate.py:613-635, 681-705), each divided by norm (the number of
 * sites for the reference's normalisation, percolate/percolate.py:1056-1060; 1.0
 * for none).  t_lo, t_hi = scipy.stats.t.interval(1 - alpha, df = runs - 1), taken
 * once by the caller.  Separately rounded IEEE operations in numpy's order: the
 * result is bit-identical to evaluating the same formulas on the host.
 *   out  host double[19 * S], S = M + 1:
 *        [0, S) runs spanning at n (k, not divided) | [S, 2S) max_cluster_size |
 *        [2S, 4S) max_cluster_size_ci[S][2] | [4S, 14S) moments_ci[5][S][2] |
 *        [14S, 19S) moments[5][S]
 */
int pz_micro_arrays(pz_ctx *ctx, double t_lo, double t_hi, double norm, double *out);

/*
 * Binomial weights: replaces _binomial_pmf (percolate/percolate.py:1067-1109),
 * one p per thread, the reference's mode-outward ratio recurrence in the
 * reference's operation order, for M bonds (no graph needed).  pmf_out: host
 * double[num_p][M+1] (may be NULL to keep the table on the device only, for
 * pz_convolve / PZ_FUSE_CANON; the latter needs M = bonds of the graph).
 */
int pz_set_ps(pz_ctx *ctx, int32_t M, int32_t num_p, const double *ps, double *pmf_out);

/*
 * out[c][p] = sum_n pmf_p[n] * cols[c][n]: the contraction of canonical_averages
 * (percolate/percolate.py:1196-1221) and of bond_canonical_statistics
 * (percolate/hpc.py:488-515) for host-resident columns.
 *   cols  host double[num_cols][M+1];  out  host double[num_cols][num_p]
 */
int pz_convolve(pz_ctx *ctx, int32_t num_cols, const double *cols, double *out);

/*
 * bond_canonical_statistics for one materialised run (percolate/hpc.py:443-515):
 *   rows  host packed rows of ONE run, (M+1) * (spanning ? 53 : 52) bytes (no graph needed);
 *   f     host double[M+1] convolution factors;  out  host double[PZ_CANON_COLS]
 *         (out[0] = percolation probability, 0 when spanning is off)
 */
int pz_canonical_statistics_rows(pz_ctx *ctx, int32_t M, int spanning, const void *rows,
                                 const double *f, double *out);

/*
 * Canonical accumulators of the fused path (PZ_FUSE_CANON), the state of
 * bond_reduce (percolate/hpc.py:638-702):
 *   count_out  runs folded;  mean_out, m2_out  host double[num_p][PZ_CANON_COLS]
 * Import merges another partial (Chan et al.) into the context -- the
 * cross-GPU step.
 */
int pz_canon_export(pz_ctx *ctx, int64_t *count_out, double *mean_out, double *m2_out);
int pz_canon_merge(pz_ctx *ctx, int64_t count, const double *mean, const double *m2);
/* forget the canonical partials only (the micro accumulators are kept) */
int pz_canon_reset(pz_ctx *ctx);

/* per-run canonical statistics of the LAST batch of runs the last pz_run_fused(PZ_FUSE_CANON)
 * call put through the device (a call is cut into batches by scratch memory; a call that fits
 * one batch -- parity tests, small R -- is covered whole): pz_canon_last_count() runs,
 * host double[count][num_p][PZ_CANON_COLS] */
int32_t pz_canon_last_count(const pz_ctx *ctx);
int pz_canon_last_runs(pz_ctx *ctx, double *out);

/*
 * Cross-GPU exchange: one process per GPU, NCCL over NVLink / NVSwitch (loaded at run time;
 * nothing here is needed on one GPU).  Replaces the reduction of the per-task results of a
 * study, bond_reduce over pickles on a shared file system (percolate/share/jugfile.py:126-135,
 * 240-244).  Rank 0 obtains an id with pz_comm_unique_id and hands its PZ_COMM_ID_BYTES bytes
 * to the other ranks by any means (a shared file, MPI, sockets); every rank then calls
 * pz_comm_init (collective).  pz_allreduce (collective) combines the accumulators of all
 * ranks' contexts in ONE step on the context's stream: the micro accumulators and the run
 * count by a word-wise integer all-reduce (exact: 32-bit limbs in 64-bit words), the canonical
 * partials by an all-gather and a Chan merge in rank order (bit-identical on every rank).
 * Afterwards every rank holds the totals.
 */
#define PZ_COMM_ID_BYTES 128
int pz_comm_unique_id(void *id_out);
int pz_comm_init(pz_ctx *ctx, int world, int rank, const void *id);
int pz_comm_destroy(pz_ctx *ctx);
int pz_comm_world(const pz_ctx *ctx);
int pz_comm_rank(const pz_ctx *ctx);
int pz_allreduce(pz_ctx *ctx);

/*
 * Per-phase device time, measured with CUDA events on the context's stream
 * around the launches of each kernel family (off by default; enabling resets
 * the counters).  ms_out / launches_out have PZ_PHASES entries.
 */
#define PZ_PHASES 7
#define PZ_PHASE_PERM 0     /* bond orders (device RNG modes)        */
#define PZ_PHASE_SWEEP 1    /* union-find sweep                      */
#define PZ_PHASE_ACCUM 2    /* per-n exact sums over runs            */
#define PZ_PHASE_CANON 3    /* per-run binomial contraction          */
#define PZ_PHASE_REDUCE 4   /* (mean, M2) over the runs of a batch   */
#define PZ_PHASE_ROWS 5     /* row materialisation                   */
#define PZ_PHASE_CKPT 6     /* run-state checkpoints (block scan)    */
int pz_profile(pz_ctx *ctx, int enable);
int pz_profile_read(pz_ctx *ctx, double *ms_out, int64_t *launches_out);

/* device-side stopwatch: CUDA events recorded on the context's stream */
int pz_timer_start(pz_ctx *ctx);
int pz_timer_stop(pz_ctx *ctx, double *ms_out);

/* number of kernels this context has launched since creation */
int64_t pz_launch_count(const pz_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* PZ_H */
