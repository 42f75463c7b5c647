// pz_api.cu -- the C-ABI of include/pz.h: context, scratch management and the
// orchestration of the kernels in pz_sweep.cu / pz_stats.cu / pz_rng.cu /
// pz_canon.cu.  No torch types; everything runs on the context's own stream.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>
#include <cmath>

#include "../../include/pz.h"
#include "pz_common.cuh"
#include "pz_internal.h"

using namespace pz;

static thread_local std::string g_err;

static int fail(int code, const std::string &msg)
{
    g_err = msg;
    return code;
}

#define PZ_CUDA(expr)                                                              \
    do {                                                                           \
        cudaError_t _e = (expr);                                                   \
        if (_e != cudaSuccess)                                                     \
            return fail(PZ_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)

template <class T>
struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;      // elements
    cudaError_t ensure(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, n * sizeof(T));
        if (e == cudaSuccess) cap = n;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct pz_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    int sms = 0;
    size_t smem_optin = 0;
    int force_kind = -1;
    int team = 1;             // CTA lock-step kernel (PZ_SWEEP_TEAM=0: single-warp A/B kernel)
    int claim_cap = 0;        // log2 of the largest claim table (PZ_CLAIM_LOG2; 0 = automatic)
    int cta_warps = 0;        // warps of a sweep CTA = bonds per batch / 32 (PZ_CTA_WARPS; 0 = automatic)
    size_t chunk_bytes = (size_t)24 << 30;      // total scratch budget of the slots (pz_create: half of
                                                // the free device memory, at most 72 GiB; PZ_CHUNK_BYTES)

    // graph
    int32_t N = 0, M = 0;
    bool spanning = false;
    int any3 = 0;
    DevBuf<uint32_t> edges32;
    DevBuf<uint2> edges64;
    DevBuf<uint32_t> sides2;

    // per-chunk scratch: PZ_SLOTS chunks of runs are in flight at once
    // (bond orders of chunk c+1 and statistics of chunk c-1 overlap the sweep
    // of chunk c on separate streams)
    struct Slot {
        DevBuf<int32_t> perms;
        DevBuf<unsigned char> recs;
        DevBuf<uint32_t> nspan;
        DevBuf<uint32_t> seeds;
        DevBuf<RunState> ckpt;
        DevBuf<double> canon_runs;        // [R][num_p][7]
        DevBuf<double> canon_red;         // 2 * num_p * 7 (mean, M2 of the chunk)
        double *red_host = nullptr;        // pinned, 2 * num_p * 7
        size_t red_host_cap = 0;
        cudaEvent_t perm_done = nullptr, sweep_done = nullptr, stats_done = nullptr;
        int32_t runs = 0;                 // runs of the chunk whose result is not yet merged
        bool canon_pending = false;
    };
    static constexpr int PZ_SLOTS = 3;
    Slot slots[PZ_SLOTS];
    cudaStream_t s_perm = nullptr, s_stats = nullptr;
    int pipeline = -1;                    // PZ_PIPELINE: 0 one stream; 1 bond orders / sweep / statistics on
                                          // three streams; 2 only the bond orders overlap the sweep;
                                          // -1 (default): 2 for the Feistel mode (no shared memory, pure ALU:
                                          // it fits next to a resident sweep CTA), else 0
    DevBuf<uint32_t> gscratch;
    DevBuf<uint8_t> rows;

    // micro accumulators
    DevBuf<unsigned long long> acc;       // (M+1) * PZ_ACC_WORDS
    DevBuf<unsigned long long> span_cum;  // M + 1
    DevBuf<double> fin;                   // 13 * (M+1): mean[7], var[6]
    DevBuf<double> arrays;                // 19 * (M+1): pz_micro_arrays
    int64_t micro_runs = 0;
    int ckpt_every = 64;                  // run state checkpoints every so many rows (64 = tile
                                          // form of accumulate; PZ_CKPT_EVERY=1024: segment form)

    // canonical
    int32_t num_p = 0;
    int32_t pmf_M = -1;                   // number of bonds the weights were built for
    int32_t sf_M = -1;                    // ... and the survival functions
    std::vector<double> ps;
    std::vector<int32_t> porder;          // sorted position -> caller's index
    DevBuf<double> ps_dev;
    DevBuf<double> pmf;                   // [num_p][M+1], rows in ascending-p order
    DevBuf<int32_t> band_lo, band_hi, tband_lo, tband_hi, porder_dev, canon_flags;
    DevBuf<double> sf;                    // survival functions [num_p][M+1]
    DevBuf<int32_t> perm_stage;           // warp-per-run shuffles: one L2-resident staging row per warp
    int gen_sms = 32;                     // SMs a warp-per-run shuffle keeps to itself next to a sweep (PZ_GEN_SMS; 0: shares them)
    DevBuf<uint32_t> validate_bits;       // caller-supplied orders: one bit per (run, bond) + flag word
    int *validate_host = nullptr;         // pinned copy of the flag word
    uint32_t epoch_start = 0x003fffffu;   // first claim epoch of a run (PZ_EPOCH_START: tests)
    bool trusted_orders = false;          // orders generated by this library: no validation
    DevBuf<double> cols;                  // scratch of the contraction
    DevBuf<double> cols_out;
    const double *canon_last_ptr = nullptr;   // per-run values of the last chunk of the last fused call
    int32_t canon_last_R = 0;
    int64_t canon_count = 0;
    std::vector<double> canon_mean, canon_m2;

    int64_t launches = 0;

    // cross-GPU exchange (NCCL, loaded at run time): see pz_comm_init / pz_allreduce
    void *comm = nullptr;                 // ncclComm_t
    int comm_world = 1, comm_rank = 0;
    DevBuf<double> comm_send, comm_recv;  // canonical partials: [1 + 2 cols] and [world][1 + 2 cols]
    double *comm_host = nullptr;          // pinned staging of both
    size_t comm_host_cap = 0;

    cudaEvent_t timer_a = nullptr, timer_b = nullptr;

    // optional per-phase device timing (CUDA events on this context's stream)
    bool profiling = false;
    double phase_ms[PZ_PHASES] = {0};
    int64_t phase_launches[PZ_PHASES] = {0};
    std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> pending;
};

struct PhaseTimer {
    pz_ctx *c; int phase; cudaStream_t st; cudaEvent_t a = nullptr, b = nullptr;
    PhaseTimer(pz_ctx *c_, int phase_, cudaStream_t st_ = nullptr)
        : c(c_), phase(phase_), st(st_ ? st_ : c_->stream) {
        if (!c->profiling) return;
        cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a, st);
    }
    ~PhaseTimer() {
        if (!c->profiling) return;
        cudaEventRecord(b, st);
        c->pending.push_back({phase, {a, b}});
    }
};

static void collect_phases(pz_ctx *c)
{
    for (auto &p : c->pending) {
        cudaEventSynchronize(p.second.second);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, p.second.first, p.second.second);
        c->phase_ms[p.first] += ms;
        c->phase_launches[p.first] += 1;
        cudaEventDestroy(p.second.first);
        cudaEventDestroy(p.second.second);
    }
    c->pending.clear();
}

// implemented in the other translation units
namespace pz {
cudaError_t launch_checkpoints(const StatsArgs &a, RunState *ckpt, int every, int n_ckpt,
                               cudaStream_t s);
cudaError_t launch_accumulate_tiles(const StatsArgs &a, unsigned long long *acc, const RunState *ckpt,
                                    int every, int n_ckpt, cudaStream_t s);
cudaError_t launch_accumulate(const StatsArgs &a, unsigned long long *acc, const RunState *ckpt,
                              int seg, int n_ckpt, cudaStream_t s);
cudaError_t launch_micro_finalize(int32_t N, int32_t M, int64_t runs, const unsigned long long *acc,
                                  unsigned long long *span_cum, double *mean, double *var,
                                  cudaStream_t s);
cudaError_t launch_micro_arrays(int32_t M, int64_t runs, double t_lo, double t_hi, double norm,
                                const double *mean, const double *var, double *out, cudaStream_t s);
cudaError_t launch_binomial_pmf(int32_t M, int32_t P, const double *ps_dev, double *pmf,
                                int32_t *xlo, int32_t *xhi, int32_t *tlo, int32_t *thi, double *sf,
                                cudaStream_t s);
cudaError_t launch_convolve(int32_t M, int32_t P, const double *pmf, const int32_t *band_lo,
                            const int32_t *band_hi, int32_t num_cols, const double *cols,
                            double *out, cudaStream_t s);
cudaError_t launch_canon_rows(int32_t M, int spanning, const uint8_t *rows, const double *f,
                              double *out, cudaStream_t s);
cudaError_t launch_canon_runs(const StatsArgs &a, int32_t P, const double *pmf, const double *sf,
                              const int32_t *xlo, const int32_t *xhi, const int32_t *tlo,
                              const int32_t *thi, const int32_t *porder, const RunState *ckpt,
                              int ckpt_every, int n_ckpt, double *out, int *flags, cudaStream_t s);
cudaError_t launch_canon_reduce(int32_t R, int32_t cols, const double *runs, double *mean,
                                double *m2, cudaStream_t s);
cudaError_t launch_perm_philox(int32_t M, int32_t R, const uint32_t *seeds, int32_t *perms,
                               cudaStream_t s, int *launches);
cudaError_t launch_perm_mt19937(int32_t M, int32_t R, const uint32_t *seeds, int32_t *perms,
                                int32_t *stage, int excl_sms, cudaStream_t s, int *launches);
size_t perm_stage_ints(int sms, int32_t M, int excl_sms);
cudaError_t launch_perm_feistel(int32_t M, int32_t R, const uint32_t *seeds, int32_t *perms,
                                cudaStream_t s, int *launches);
cudaError_t launch_perm_philox_fy(int32_t M, int32_t R, const uint32_t *seeds, int32_t *perms,
                                  int32_t *stage, int excl_sms, cudaStream_t s, int *launches);
cudaError_t launch_validate_orders(int32_t M, int32_t R, const int32_t *perms, uint32_t *bitmap,
                                   int *flag, cudaStream_t s);
}

static cudaError_t launch_perm_mode(pz_ctx *c, int perm_mode, int32_t M, int32_t R, const uint32_t *seeds,
                                    int32_t *perms, cudaStream_t s, int *l, int excl_sms = 0)
{
    switch (perm_mode) {
    case PZ_PERM_PHILOX: return launch_perm_philox(M, R, seeds, perms, s, l);
    case PZ_PERM_FEISTEL: return launch_perm_feistel(M, R, seeds, perms, s, l);
    default: break;
    }
    // warp-per-run shuffles: one staging row per warp of the launch (allocated once per graph size;
    // launches of one context are ordered on a stream, so they never share it concurrently)
    cudaError_t e = c->perm_stage.ensure(perm_stage_ints(c->sms, M, c->gen_sms));
    if (e != cudaSuccess) return e;
    if (perm_mode == PZ_PERM_PHILOX_FY)
        return launch_perm_philox_fy(M, R, seeds, perms, c->perm_stage.p, excl_sms, s, l);
    return launch_perm_mt19937(M, R, seeds, perms, c->perm_stage.p, excl_sms, s, l);
}

extern "C" {

const char *pz_last_error(void) { return g_err.c_str(); }
int pz_version(void) { return 100; }

int pz_create(int device, pz_ctx **out)
{
    if (!out) return fail(PZ_ERR_ARG, "pz_create: out is NULL");
    int count = 0;
    PZ_CUDA(cudaGetDeviceCount(&count));
    if (device < 0 || device >= count)
        return fail(PZ_ERR_ARG, "pz_create: no such CUDA device");
    PZ_CUDA(cudaSetDevice(device));
    pz_ctx *c = new pz_ctx();
    c->device = device;
    cudaDeviceProp prop;
    PZ_CUDA(cudaGetDeviceProperties(&prop, device));
    c->sms = prop.multiProcessorCount;
    c->smem_optin = prop.sharedMemPerBlockOptin;
    PZ_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    PZ_CUDA(cudaStreamCreateWithFlags(&c->s_perm, cudaStreamNonBlocking));
    PZ_CUDA(cudaStreamCreateWithFlags(&c->s_stats, cudaStreamNonBlocking));
    for (auto &sl : c->slots) {
        PZ_CUDA(cudaEventCreateWithFlags(&sl.perm_done, cudaEventDisableTiming));
        PZ_CUDA(cudaEventCreateWithFlags(&sl.sweep_done, cudaEventDisableTiming));
        PZ_CUDA(cudaEventCreateWithFlags(&sl.stats_done, cudaEventDisableTiming));
    }
    if (const char *e = getenv("PZ_PIPELINE")) c->pipeline = atoi(e);
    if (const char *e = getenv("PZ_GEN_SMS")) c->gen_sms = std::max(0, std::min(atoi(e), 100));
    if (const char *e = getenv("PZ_FORCE_STORE")) c->force_kind = atoi(e);
    if (const char *e = getenv("PZ_SWEEP_TEAM")) c->team = atoi(e);
    if (const char *e = getenv("PZ_CLAIM_LOG2")) c->claim_cap = atoi(e);
    if (const char *e = getenv("PZ_CTA_WARPS")) c->cta_warps = atoi(e);
    if (const char *e = getenv("PZ_CKPT_EVERY")) c->ckpt_every = atoi(e) == 64 ? 64 : 1024;
    if (const char *e = getenv("PZ_EPOCH_START")) {
        const long v = atol(e);
        c->epoch_start = (uint32_t)std::min<long>(0x003fffffL, std::max<long>(0x1280L, v));
    }
    PZ_CUDA(cudaMallocHost(&c->validate_host, sizeof(int)));
    {
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && free_b > 0)
            c->chunk_bytes = std::min<size_t>(free_b / 2, (size_t)72 << 30);
    }
    if (const char *e = getenv("PZ_CHUNK_BYTES")) c->chunk_bytes = (size_t)atoll(e);
    *out = c;
    return PZ_OK;
}

void pz_destroy(pz_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    c->edges32.release(); c->edges64.release(); c->sides2.release();
    for (auto &sl : c->slots) {
        sl.perms.release(); sl.recs.release(); sl.nspan.release(); sl.seeds.release();
        sl.ckpt.release(); sl.canon_runs.release(); sl.canon_red.release();
        if (sl.red_host) cudaFreeHost(sl.red_host);
        if (sl.perm_done) cudaEventDestroy(sl.perm_done);
        if (sl.sweep_done) cudaEventDestroy(sl.sweep_done);
        if (sl.stats_done) cudaEventDestroy(sl.stats_done);
    }
    c->validate_bits.release();
    c->perm_stage.release();
    if (c->validate_host) cudaFreeHost(c->validate_host);
    pz_comm_destroy(c);
    c->comm_send.release(); c->comm_recv.release();
    if (c->comm_host) cudaFreeHost(c->comm_host);
    c->gscratch.release(); c->rows.release(); c->acc.release(); c->span_cum.release();
    c->fin.release(); c->arrays.release(); c->ps_dev.release(); c->band_lo.release();
    c->band_hi.release(); c->tband_lo.release(); c->tband_hi.release(); c->canon_flags.release();
    c->sf.release(); c->porder_dev.release(); c->cols.release(); c->cols_out.release();
    c->pmf.release();
    if (c->timer_a) { cudaEventDestroy(c->timer_a); cudaEventDestroy(c->timer_b); }
    cudaStreamDestroy(c->s_perm);
    cudaStreamDestroy(c->s_stats);
    cudaStreamDestroy(c->stream);
    delete c;
}

int pz_device(const pz_ctx *c) { return c ? c->device : -1; }
void *pz_stream(const pz_ctx *c) { return c ? (void *)c->stream : nullptr; }
int64_t pz_launch_count(const pz_ctx *c) { return c ? c->launches : 0; }

int pz_synchronize(pz_ctx *c)
{
    if (!c) return fail(PZ_ERR_ARG, "null context");
    PZ_CUDA(cudaSetDevice(c->device));
    PZ_CUDA(cudaStreamSynchronize(c->s_perm));
    PZ_CUDA(cudaStreamSynchronize(c->stream));
    PZ_CUDA(cudaStreamSynchronize(c->s_stats));
    return PZ_OK;
}

int pz_set_graph(pz_ctx *c, int32_t N, int32_t M, const int32_t *eu, const int32_t *ev,
                 const uint8_t *side_mask, int preconnected)
{
    if (!c) return fail(PZ_ERR_ARG, "null context");
    if (N < 1 || M < 0 || (M > 0 && (!eu || !ev)))
        return fail(PZ_ERR_ARG, "pz_set_graph: need N >= 1, M >= 0 and endpoint arrays");
    if (N > (1 << 29)) return fail(PZ_ERR_ARG, "pz_set_graph: N must be <= 2^29");
    for (int32_t e = 0; e < M; ++e)
        if (eu[e] < 0 || eu[e] >= N || ev[e] < 0 || ev[e] >= N)
            return fail(PZ_ERR_ARG, "pz_set_graph: bond endpoint out of range");
    PZ_CUDA(cudaSetDevice(c->device));
    PZ_CUDA(cudaStreamSynchronize(c->stream));
    c->N = N; c->M = M;
    c->spanning = side_mask != nullptr;
    c->any3 = 0;

    // Global-memory store (N > 65536): the union-find array lives in L2 / HBM and every node visit
    // is a 32-byte sector.  Node ids are internal (rows carry bond indices only), so nodes are
    // renumbered by small breadth-first neighbourhoods: the two endpoints of a bond, and a node and
    // the root of its (small) cluster, then usually share a sector or a line.
    std::vector<int32_t> relabel;                      // old id -> new id (empty = identity)
    {
        int blob = (N > 65536 || c->force_kind == STORE_G32) ? 32 : 0;
        if (const char *e = getenv("PZ_RELABEL")) blob = atoi(e);
        if (blob > 1 && M > 0) {
            std::vector<int64_t> off((size_t)N + 1, 0);
            for (int32_t e = 0; e < M; ++e) { ++off[(size_t)eu[e] + 1]; ++off[(size_t)ev[e] + 1]; }
            for (int32_t x = 0; x < N; ++x) off[(size_t)x + 1] += off[x];
            std::vector<int32_t> adj((size_t)2 * M);
            {
                std::vector<int64_t> fill(off.begin(), off.end() - 1);
                for (int32_t e = 0; e < M; ++e) { adj[fill[eu[e]]++] = ev[e]; adj[fill[ev[e]]++] = eu[e]; }
            }
            relabel.assign((size_t)N, -1);
            std::vector<int32_t> q;
            q.reserve((size_t)blob);
            int32_t next = 0;
            for (int32_t s0 = 0; s0 < N; ++s0) {
                if (relabel[s0] >= 0) continue;
                q.clear();
                q.push_back(s0);
                relabel[s0] = next++;
                for (size_t head = 0; head < q.size() && (int)q.size() < blob; ++head) {
                    const int32_t x = q[head];
                    for (int64_t k = off[x]; k < off[(size_t)x + 1] && (int)q.size() < blob; ++k) {
                        const int32_t y = adj[k];
                        if (relabel[y] < 0) { relabel[y] = next++; q.push_back(y); }
                    }
                }
            }
        }
    }
    auto id = [&](int32_t x) -> uint32_t { return (uint32_t)(relabel.empty() ? x : relabel[x]); };

    std::vector<uint2> e64((size_t)std::max(M, 1));
    for (int32_t e = 0; e < M; ++e) e64[e] = make_uint2(id(eu[e]), id(ev[e]));
    PZ_CUDA(c->edges64.ensure(e64.size()));
    PZ_CUDA(cudaMemcpyAsync(c->edges64.p, e64.data(), e64.size() * sizeof(uint2),
                            cudaMemcpyHostToDevice, c->stream));
    std::vector<uint32_t> e32;
    if (N <= 65536) {
        e32.resize((size_t)std::max(M, 1));
        for (int32_t e = 0; e < M; ++e) e32[e] = id(eu[e]) | (id(ev[e]) << 16);
        PZ_CUDA(c->edges32.ensure(e32.size()));
        PZ_CUDA(cudaMemcpyAsync(c->edges32.p, e32.data(), e32.size() * 4,
                                cudaMemcpyHostToDevice, c->stream));
    }
    std::vector<uint32_t> s2;
    if (side_mask) {
        s2.assign((size_t)(N + 15) / 16, 0u);
        int any3 = preconnected ? 1 : 0;
        for (int32_t x = 0; x < N; ++x) {
            const uint32_t m = side_mask[x] & 3u;
            if (m == 3u) any3 = 1;
            const uint32_t nx = id(x);
            s2[nx >> 4] |= m << ((nx & 15) * 2);
        }
        c->any3 = any3;
        PZ_CUDA(c->sides2.ensure(s2.size()));
        PZ_CUDA(cudaMemcpyAsync(c->sides2.p, s2.data(), s2.size() * 4,
                                cudaMemcpyHostToDevice, c->stream));
    }
    PZ_CUDA(cudaStreamSynchronize(c->stream));
    // graph change invalidates everything derived from it
    c->micro_runs = 0;
    c->acc.release(); c->span_cum.release();
    c->canon_count = 0; c->canon_last_R = 0;
    return PZ_OK;
}

int pz_row_bytes(const pz_ctx *c) { return c ? (c->spanning ? 53 : 52) : 0; }

}  // extern "C"

// ---------------------------------------------------------------------------
// one chunk of runs: bond orders on the device, then the sweep
// ---------------------------------------------------------------------------
struct Chunk {
    SweepPlan plan;
    StatsArgs stats;
    const int32_t *perms_dev;
};

// bond orders of the chunk on `sp` (slot buffers), then the sweep on the
// context's main stream; `sl.sweep_done` is recorded behind the sweep
static int sweep_chunk(pz_ctx *c, pz_ctx::Slot &sl, cudaStream_t sp, int32_t R, int perm_mode_in,
                       const void *perm_src, size_t run0, Chunk *out, int gen_excl_sms = 0,
                       int sweep_grid_cap = 0)
{
    const int32_t M = c->M;
    const bool seeds_on_device = (perm_mode_in & PZ_SEEDS_ON_DEVICE) != 0;
    const int perm_mode = perm_mode_in & ~PZ_SEEDS_ON_DEVICE;
    const int32_t *perms_dev = nullptr;
    const size_t pm = (size_t)R * (size_t)std::max(M, 1);
    if (perm_mode == PZ_PERM_DEVICE) {
        perms_dev = (const int32_t *)perm_src + run0 * (size_t)M;
    } else {
        PZ_CUDA(sl.perms.ensure(pm));
        perms_dev = sl.perms.p;
        if (perm_mode == PZ_PERM_HOST) {
            if (M > 0)
                PZ_CUDA(cudaMemcpyAsync(sl.perms.p, (const int32_t *)perm_src + run0 * (size_t)M,
                                        (size_t)R * M * 4, cudaMemcpyHostToDevice, sp));
        } else {
            const uint32_t *seeds_dev = (const uint32_t *)perm_src + run0;
            if (!seeds_on_device) {
                PZ_CUDA(sl.seeds.ensure((size_t)R));
                PZ_CUDA(cudaMemcpyAsync(sl.seeds.p, (const uint32_t *)perm_src + run0, (size_t)R * 4,
                                        cudaMemcpyHostToDevice, sp));
                seeds_dev = sl.seeds.p;
            }
            int l = 0;
            PhaseTimer t(c, PZ_PHASE_PERM, sp);
            PZ_CUDA(launch_perm_mode(c, perm_mode, M, R, seeds_dev, sl.perms.p, sp, &l, gen_excl_sms));
            c->launches += l;
        }
    }
    if ((perm_mode == PZ_PERM_HOST || perm_mode == PZ_PERM_DEVICE) && !c->trusted_orders && M > 0 && R > 0) {
        // caller-supplied orders: range and permutation check before anything indexes with them
        const size_t words = (size_t)R * (((size_t)M + 31) / 32) + 1;
        PZ_CUDA(c->validate_bits.ensure(words));
        int *flag = reinterpret_cast<int *>(c->validate_bits.p + (words - 1));
        PZ_CUDA(launch_validate_orders(M, R, perms_dev, c->validate_bits.p, flag, sp));
        c->launches += 1;
        PZ_CUDA(cudaMemcpyAsync(c->validate_host, flag, sizeof(int), cudaMemcpyDeviceToHost, sp));
        PZ_CUDA(cudaStreamSynchronize(sp));
        if (*c->validate_host & 1)
            return fail(PZ_ERR_ARG, "bond order entry outside [0, num_edges)");
        if (*c->validate_host & 2)
            return fail(PZ_ERR_ARG, "bond order is not a permutation (repeated entry)");
    }
    SweepPlan plan = plan_sweep(c->N, R, c->sms, c->smem_optin, c->force_kind, c->team, c->claim_cap, c->cta_warps);
    if (sweep_grid_cap > 0 && plan.grid > sweep_grid_cap) plan.grid = sweep_grid_cap;   // (runs are grid-stride)
    if (plan.kind != STORE_G32 && c->N > 65536)
        return fail(PZ_ERR_ARG, "forced shared-memory store needs N <= 65536");
    const bool rec64 = plan.kind == STORE_G32;
    PZ_CUDA(sl.recs.ensure(pm * (rec64 ? 8 : 4)));
    PZ_CUDA(sl.nspan.ensure((size_t)R));
    if (plan.gscratch_bytes) PZ_CUDA(c->gscratch.ensure(plan.gscratch_bytes / 4));
    if (sp != c->stream) {
        PZ_CUDA(cudaEventRecord(sl.perm_done, sp));
        PZ_CUDA(cudaStreamWaitEvent(c->stream, sl.perm_done, 0));
    }

    SweepArgs sa{};
    sa.N = c->N; sa.M = M; sa.R = R;
    sa.edges = rec64 ? (const void *)c->edges64.p : (const void *)c->edges32.p;
    sa.sides2 = c->spanning ? c->sides2.p : nullptr;
    sa.any3 = c->any3;
    sa.perms = perms_dev;
    sa.recs = sl.recs.p;
    sa.nspan = sl.nspan.p;
    sa.gscratch = c->gscratch.p;
    sa.claim_log2 = plan.claim_log2;
    sa.epoch_start = c->epoch_start;
    {
        PhaseTimer t(c, PZ_PHASE_SWEEP);
        PZ_CUDA(launch_sweep(plan, sa, c->stream));
    }
    c->launches += 1;
    PZ_CUDA(cudaEventRecord(sl.sweep_done, c->stream));

    out->plan = plan;
    out->perms_dev = perms_dev;
    out->stats = StatsArgs{c->N, M, R, rec64 ? 1 : 0, sl.recs.p, sl.nspan.p, perms_dev,
                           c->spanning ? 1 : 0};
    return PZ_OK;
}

static int check_run_args(pz_ctx *c, int32_t R, int perm_mode, const void *perm_src)
{
    if (!c) return fail(PZ_ERR_ARG, "null context");
    if (c->N == 0) return fail(PZ_ERR_STATE, "no graph set (call pz_set_graph first)");
    if (R < 0) return fail(PZ_ERR_ARG, "R must be >= 0");
    const int base_mode = perm_mode & ~PZ_SEEDS_ON_DEVICE;
    if (base_mode < PZ_PERM_HOST || base_mode > PZ_PERM_PHILOX_FY)
        return fail(PZ_ERR_ARG, "unknown perm_mode");
    if ((perm_mode & PZ_SEEDS_ON_DEVICE) && base_mode < PZ_PERM_MT19937)
        return fail(PZ_ERR_ARG, "PZ_SEEDS_ON_DEVICE needs a device RNG mode");
    if (R > 0 && !perm_src && c->M > 0) return fail(PZ_ERR_ARG, "perm_src is NULL");
    if (base_mode == PZ_PERM_PHILOX && (long long)c->M > (1ll << 24))
        return fail(PZ_ERR_ARG, "PZ_PERM_PHILOX supports at most 2^24 bonds (use PZ_PERM_PHILOX_FY or PZ_PERM_MT19937)");
    return PZ_OK;
}

extern "C" {

int pz_run_rows(pz_ctx *c, int32_t R, int perm_mode, const void *perm_src, void *rows_out,
                int32_t *perms_out)
{
    int rc = check_run_args(c, R, perm_mode, perm_src);
    if (rc) return rc;
    if (R > 0 && !rows_out) return fail(PZ_ERR_ARG, "rows_out is NULL");
    PZ_CUDA(cudaSetDevice(c->device));
    const size_t rb = c->spanning ? 53 : 52;
    const size_t run_bytes = ((size_t)c->M + 1) * rb;
    size_t chunk = std::max<size_t>(1, (c->chunk_bytes / pz_ctx::PZ_SLOTS) / (run_bytes + 8 * (size_t)c->M));
    for (size_t r0 = 0; r0 < (size_t)R; r0 += chunk) {
        const int32_t rc_n = (int32_t)std::min(chunk, (size_t)R - r0);
        Chunk ch;
        rc = sweep_chunk(c, c->slots[0], c->stream, rc_n, perm_mode, perm_src, r0, &ch);
        if (rc) return rc;
        PZ_CUDA(c->rows.ensure((size_t)rc_n * run_bytes));
        {
            PhaseTimer t(c, PZ_PHASE_ROWS);
            PZ_CUDA(launch_expand_rows(ch.stats, c->rows.p, c->stream));
        }
        c->launches += 1;
        PZ_CUDA(cudaMemcpyAsync((uint8_t *)rows_out + r0 * run_bytes, c->rows.p,
                                (size_t)rc_n * run_bytes, cudaMemcpyDeviceToHost, c->stream));
        if (perms_out && c->M > 0)
            PZ_CUDA(cudaMemcpyAsync(perms_out + r0 * (size_t)c->M, ch.perms_dev,
                                    (size_t)rc_n * c->M * 4, cudaMemcpyDeviceToHost, c->stream));
        PZ_CUDA(cudaStreamSynchronize(c->stream));
    }
    return PZ_OK;
}

}  // extern "C"


// ---------------------------------------------------------------------------
// fused path
// ---------------------------------------------------------------------------
static int ensure_acc(pz_ctx *c)
{
    // (+ one word behind the block: the run count travels with the cross-GPU all-reduce)
    const size_t words = ((size_t)c->M + 1) * PZ_ACC_WORDS + 1;
    if (c->acc.cap < words) {
        PZ_CUDA(c->acc.ensure(words));
        PZ_CUDA(cudaMemsetAsync(c->acc.p, 0, words * 8, c->stream));
        c->micro_runs = 0;
    }
    return PZ_OK;
}

static void chan_merge(int64_t &na, std::vector<double> &mean_a, std::vector<double> &m2_a,
                       int64_t nb, const double *mean_b, const double *m2_b)
{
    // Chan et al. pairwise merge -- the arithmetic bond_reduce delegates to
    // simoa.stats.online_variance (percolate/hpc.py:677-684)
    if (nb <= 0) return;
    if (na == 0) {
        mean_a.assign(mean_b, mean_b + mean_a.size());
        m2_a.assign(m2_b, m2_b + m2_a.size());
        na = nb;
        return;
    }
    const double fa = (double)na, fb = (double)nb, n = fa + fb;
    for (size_t i = 0; i < mean_a.size(); ++i) {
        const double delta = mean_b[i] - mean_a[i];
        mean_a[i] = mean_a[i] + delta * fb / n;
        m2_a[i] = m2_a[i] + m2_b[i] + delta * delta * fa * fb / n;
    }
    na += nb;
}

extern "C" {

int pz_make_perms(pz_ctx *c, int32_t R, int perm_mode, const uint32_t *seeds, int32_t *out,
                  int is_device)
{
    if (!c) return fail(PZ_ERR_ARG, "null context");
    if (c->N == 0) return fail(PZ_ERR_STATE, "no graph set (call pz_set_graph first)");
    if (perm_mode < PZ_PERM_MT19937 || perm_mode > PZ_PERM_PHILOX_FY)
        return fail(PZ_ERR_ARG, "pz_make_perms: perm_mode must be a device RNG mode");
    if (R < 0 || (R > 0 && (!seeds || !out))) return fail(PZ_ERR_ARG, "pz_make_perms: bad arguments");
    if (R == 0 || c->M == 0) return PZ_OK;
    if (perm_mode == PZ_PERM_PHILOX && (long long)c->M > (1ll << 24))
        return fail(PZ_ERR_ARG, "PZ_PERM_PHILOX supports at most 2^24 bonds (use PZ_PERM_PHILOX_FY or PZ_PERM_MT19937)");
    PZ_CUDA(cudaSetDevice(c->device));
    const size_t per_run = (size_t)c->M * 4;
    const size_t chunk = is_device ? (size_t)R
                                   : std::max<size_t>(1, (c->chunk_bytes / pz_ctx::PZ_SLOTS) / per_run);
    for (size_t r0 = 0; r0 < (size_t)R; r0 += chunk) {
        const int32_t n = (int32_t)std::min(chunk, (size_t)R - r0);
        pz_ctx::Slot &sl = c->slots[0];
        PZ_CUDA(sl.seeds.ensure((size_t)n));
        PZ_CUDA(cudaMemcpyAsync(sl.seeds.p, seeds + r0, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
        int32_t *dst = out + r0 * (size_t)c->M;
        if (!is_device) { PZ_CUDA(sl.perms.ensure((size_t)n * c->M)); dst = sl.perms.p; }
        int l = 0;
        {
            PhaseTimer t(c, PZ_PHASE_PERM);
            PZ_CUDA(launch_perm_mode(c, perm_mode, c->M, n, sl.seeds.p, dst, c->stream, &l));
        }
        c->launches += l;
        if (!is_device)
            PZ_CUDA(cudaMemcpyAsync(out + r0 * (size_t)c->M, dst, (size_t)n * per_run,
                                    cudaMemcpyDeviceToHost, c->stream));
        PZ_CUDA(cudaStreamSynchronize(c->stream));
        if (c->profiling) collect_phases(c);
    }
    return PZ_OK;
}

int pz_reset_accumulators(pz_ctx *c)
{
    if (!c) return fail(PZ_ERR_ARG, "null context");
    PZ_CUDA(cudaSetDevice(c->device));
    if (c->acc.p)
        PZ_CUDA(cudaMemsetAsync(c->acc.p, 0, ((size_t)c->M + 1) * PZ_ACC_WORDS * 8, c->stream));
    c->micro_runs = 0;
    c->canon_count = 0;
    c->canon_last_R = 0;
    return PZ_OK;
}

// fold the (mean, M2) of a finished chunk into the context (chunks are
// harvested in submission order, so the result does not depend on timing)
static int harvest_slot(pz_ctx *c, pz_ctx::Slot &sl)
{
    if (!sl.canon_pending) return PZ_OK;
    PZ_CUDA(cudaEventSynchronize(sl.stats_done));
    const size_t cols = (size_t)c->num_p * PZ_CANON_COLS;
    c->canon_mean.resize(cols); c->canon_m2.resize(cols);
    chan_merge(c->canon_count, c->canon_mean, c->canon_m2, sl.runs, sl.red_host,
               sl.red_host + cols);
    sl.canon_pending = false;
    return PZ_OK;
}

int pz_run_fused(pz_ctx *c, int32_t R, int perm_mode, const void *perm_src, int flags)
{
    int rc = check_run_args(c, R, perm_mode, perm_src);
    if (rc) return rc;
    if (!(flags & (PZ_FUSE_MICRO | PZ_FUSE_CANON))) return fail(PZ_ERR_ARG, "pz_run_fused: no flags");
    if ((flags & PZ_FUSE_CANON) && (c->num_p == 0 || c->pmf_M != c->M || c->sf_M != c->M))
        return fail(PZ_ERR_STATE, "pz_run_fused: PZ_FUSE_CANON needs pz_set_ps(M = bonds of the graph) first");
    PZ_CUDA(cudaSetDevice(c->device));
    const int base_mode = perm_mode & ~PZ_SEEDS_ON_DEVICE;
    if (flags & PZ_FUSE_MICRO) { rc = ensure_acc(c); if (rc) return rc; }
    // everything issued so far on the main stream (graph, weights, resets) must
    // be visible to the side streams
    PZ_CUDA(cudaStreamSynchronize(c->stream));
    const int P = c->num_p;
    const SweepPlan p0 = plan_sweep(c->N, R, c->sms, c->smem_optin, c->force_kind, c->team,
                                    c->claim_cap, c->cta_warps);
    // Streams (measured, profiles/rng_r2.txt).  2: the bond orders of chunk k+1 on a second stream
    // under the sweep of chunk k.  The bucketed Philox shuffle needs shared memory: next to a sweep
    // that fills the SMs' shared memory its CTAs follow the sweep's as those retire and run beside
    // the statistics kernels (-1 %); next to a global-memory sweep they would share the SMs with it
    // and halve its pace (L = 1024: 420 instead of 284 ms), so there it stays on the main stream.
    // 1: in addition the statistics kernels of chunk k on a third stream under the sweep of chunk
    // k+1: +4.5 % at 12500 runs, +2.4 % at 25000, +0.5 % at 50000, -3 % at 100000 (the contraction
    // kernel crawls next to a long sweep), hence for calls of moderate size only.
    const int pipeline = (base_mode == PZ_PERM_HOST || base_mode == PZ_PERM_DEVICE) ? 0
                         : c->pipeline >= 0 ? c->pipeline
                         : (base_mode == PZ_PERM_PHILOX && p0.kind == STORE_G32) ? 0
                         : (p0.kind != STORE_G32 && R <= 60000) ? 1 : 2;
    // three slots rotate even on one stream: the host then runs up to two chunks ahead of the
    // device instead of waiting for every chunk's statistics before it launches the next sweep
    const int nslot = pz_ctx::PZ_SLOTS;
    const bool split_sms = (base_mode == PZ_PERM_MT19937 || base_mode == PZ_PERM_PHILOX_FY) && c->gen_sms > 0 &&
                           p0.kind != STORE_G32 && p0.grid >= c->sms && p0.smem_bytes > (size_t)114 * 1024;
    const bool hide = pipeline != 0;          // bond orders are generated underneath the previous sweep
    const size_t per_run = (size_t)std::max(c->M, 1) * 12 + 4096 +   // orders + widest records
                           ((size_t)c->M / c->ckpt_every + 1) * 32;  // + checkpoints
    // the global-memory store keeps one parent array per resident sweep CTA next to the slots
    size_t budget = c->chunk_bytes;
    budget = p0.gscratch_bytes < budget / 2 ? budget - p0.gscratch_bytes : budget / 2;
    size_t chunk = std::max<size_t>(1, (budget / pz_ctx::PZ_SLOTS) / per_run);
    // Chunk sizes: whole waves (one run per sweep CTA, grid = a multiple of the SM count), as few
    // equal chunks as the scratch allows.  With the bond orders on their own stream: at least three
    // chunks (all but the first chunk's orders are generated underneath a sweep; never fewer than
    // 16 waves per chunk, small chunks starve the statistics kernels), and when the call is long
    // anyway (>= 3 chunks by memory) a short ramp in front (4 waves, then x3) so that the orders
    // generated with nothing to hide behind are those of a small chunk.
    // Measured (one B200, L = 256): 12500 runs 127.7 ms with three equal chunks, 131.3 ms with the
    // ramp; 100000 runs 980 ms without the ramp, 972 ms with it.
    std::vector<size_t> sizes;
    {
        const size_t sms = (size_t)std::max(c->sms, 1);
        size_t rem = (size_t)R;
        size_t k = std::max<size_t>(1, (rem + chunk - 1) / chunk);
        if (hide && k >= 3) {
            size_t last = 0;
            for (size_t ramp = 4 * sms; ramp < chunk && rem > 4 * ramp; ramp *= 3) {
                sizes.push_back(ramp);
                rem -= ramp;
                last = ramp;
            }
            k = std::max<size_t>(1, (rem + chunk - 1) / chunk);
            // (generating the orders of a run takes about a fifth of sweeping it: a chunk of up to
            // four times the previous one still hides its orders)
            if (last) k = std::max(k, (rem + 4 * last - 1) / (4 * last));
        } else if (hide) {
            k = std::max<size_t>(k, std::min<size_t>(3, rem / (16 * sms)));
            k = std::max<size_t>(k, 1);
        }
        size_t even = (rem + k - 1) / k;
        if (k > 1 && even > sms && chunk >= sms)
            even = std::min(chunk - chunk % sms, (even + sms - 1) / sms * sms);
        even = std::max<size_t>(std::min(even, chunk), 1);
        for (; rem > 0; rem -= std::min(rem, even)) sizes.push_back(std::min(rem, even));
    }
    const int n_ckpt = c->M / c->ckpt_every + 1;
    cudaStream_t sp = pipeline ? c->s_perm : c->stream;
    cudaStream_t ss = pipeline == 1 ? c->s_stats : c->stream;    // 2: only the bond orders overlap
    size_t ci = 0;
    for (size_t r0 = 0; ci < sizes.size(); r0 += sizes[ci], ++ci) {
        const int32_t n = (int32_t)sizes[ci];
        pz_ctx::Slot &sl = c->slots[ci % nslot];
        // the slot's previous chunk must be done before its buffers are reused
        if (ci >= (size_t)nslot) {
            rc = harvest_slot(c, sl); if (rc) return rc;
            PZ_CUDA(cudaEventSynchronize(sl.stats_done));
        }
        Chunk ch;
        // Warp-per-run shuffles next to a one-CTA-per-SM sweep: sharing the SMs costs the sweep 40 %
        // (the shuffle's uncoalesced accesses share the load/store pipe with the sweep's shared-
        // memory chains), so the bond orders of chunk ci are generated on gen_sms SMs of their own
        // while the sweep of chunk ci - 1 runs on the others (its grid is capped accordingly; the
        // first chunk's orders have the whole GPU, the last sweep as well).
        const bool part = split_sms && pipeline != 0;
        rc = sweep_chunk(c, sl, sp, n, perm_mode, perm_src, r0, &ch, part && ci > 0 ? c->gen_sms : 0,
                         part && ci + 1 < sizes.size() ? c->sms - c->gen_sms : 0);
        if (rc) return rc;
        PZ_CUDA(sl.ckpt.ensure((size_t)n * n_ckpt));
        if (ss != c->stream) PZ_CUDA(cudaStreamWaitEvent(ss, sl.sweep_done, 0));
        {
            PhaseTimer t(c, PZ_PHASE_CKPT, ss);
            PZ_CUDA(launch_checkpoints(ch.stats, sl.ckpt.p, c->ckpt_every, n_ckpt, ss));
        }
        c->launches += 1;
        if (flags & PZ_FUSE_MICRO) {
            PhaseTimer t(c, PZ_PHASE_ACCUM, ss);
            if (c->ckpt_every == 64)
                PZ_CUDA(launch_accumulate_tiles(ch.stats, c->acc.p, sl.ckpt.p, c->ckpt_every, n_ckpt, ss));
            else
                PZ_CUDA(launch_accumulate(ch.stats, c->acc.p, sl.ckpt.p, c->ckpt_every, n_ckpt, ss));
            c->launches += 1;
            c->micro_runs += n;
        }
        if (flags & PZ_FUSE_CANON) {
            const int cols = P * PZ_CANON_COLS;
            PZ_CUDA(sl.canon_runs.ensure((size_t)n * cols));
            PZ_CUDA(sl.canon_red.ensure((size_t)2 * cols));
            {
                PhaseTimer t(c, PZ_PHASE_CANON, ss);
                PZ_CUDA(launch_canon_runs(ch.stats, P, c->pmf.p, c->sf.p, c->band_lo.p, c->band_hi.p,
                                          c->tband_lo.p, c->tband_hi.p, c->porder_dev.p, sl.ckpt.p,
                                          c->ckpt_every, n_ckpt, sl.canon_runs.p, c->canon_flags.p,
                                          ss));
            }
            {
                PhaseTimer t(c, PZ_PHASE_REDUCE, ss);
                PZ_CUDA(launch_canon_reduce(n, cols, sl.canon_runs.p, sl.canon_red.p,
                                            sl.canon_red.p + cols, ss));
            }
            c->launches += 3;
            if (sl.red_host_cap < (size_t)2 * cols) {
                if (sl.red_host) cudaFreeHost(sl.red_host);
                sl.red_host = nullptr; sl.red_host_cap = 0;
                PZ_CUDA(cudaMallocHost(&sl.red_host, (size_t)2 * cols * 8));
                sl.red_host_cap = (size_t)2 * cols;
            }
            PZ_CUDA(cudaMemcpyAsync(sl.red_host, sl.canon_red.p, (size_t)2 * cols * 8,
                                    cudaMemcpyDeviceToHost, ss));
            sl.runs = n;
            sl.canon_pending = true;
            c->canon_last_R = n;
            c->canon_last_ptr = sl.canon_runs.p;
        }
        PZ_CUDA(cudaEventRecord(sl.stats_done, ss));
    }
    // drain: harvest the remaining chunks in submission order
    for (size_t k = (ci > (size_t)nslot ? ci - nslot : 0); k < ci; ++k) {
        rc = harvest_slot(c, c->slots[k % nslot]); if (rc) return rc;
    }
    PZ_CUDA(cudaStreamSynchronize(sp));
    PZ_CUDA(cudaStreamSynchronize(c->stream));
    PZ_CUDA(cudaStreamSynchronize(ss));
    if (c->profiling) collect_phases(c);
    return PZ_OK;
}

int64_t pz_micro_runs(const pz_ctx *c) { return c ? c->micro_runs : 0; }

int pz_micro_export(pz_ctx *c, uint64_t *dst, int is_device)
{
    if (!c || !dst) return fail(PZ_ERR_ARG, "pz_micro_export: bad arguments");
    if (c->N == 0) return fail(PZ_ERR_STATE, "no graph set");
    PZ_CUDA(cudaSetDevice(c->device));
    int rc = ensure_acc(c); if (rc) return rc;
    PZ_CUDA(cudaMemcpyAsync(dst, c->acc.p, ((size_t)c->M + 1) * PZ_ACC_WORDS * 8,
                            is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->stream));
    PZ_CUDA(cudaStreamSynchronize(c->stream));
    return PZ_OK;
}

int pz_micro_import(pz_ctx *c, const uint64_t *src, int is_device, int64_t runs)
{
    if (!c || !src || runs < 0) return fail(PZ_ERR_ARG, "pz_micro_import: bad arguments");
    if (c->N == 0) return fail(PZ_ERR_STATE, "no graph set");
    PZ_CUDA(cudaSetDevice(c->device));
    int rc = ensure_acc(c); if (rc) return rc;
    PZ_CUDA(cudaMemcpyAsync(c->acc.p, src, ((size_t)c->M + 1) * PZ_ACC_WORDS * 8,
                            is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, c->stream));
    PZ_CUDA(cudaStreamSynchronize(c->stream));
    c->micro_runs = runs;
    return PZ_OK;
}

int pz_micro_finalize(pz_ctx *c, double *mean_out, double *var_out)
{
    if (!c || (!mean_out) != (!var_out)) return fail(PZ_ERR_ARG, "pz_micro_finalize: bad arguments");
    if (c->N == 0) return fail(PZ_ERR_STATE, "no graph set");
    if (c->micro_runs <= 0 || !c->acc.p) return fail(PZ_ERR_STATE, "pz_micro_finalize: no runs accumulated");
    PZ_CUDA(cudaSetDevice(c->device));
    const size_t S = (size_t)c->M + 1;
    PZ_CUDA(c->span_cum.ensure(S));
    PZ_CUDA(c->fin.ensure(13 * S));
    PZ_CUDA(launch_micro_finalize(c->N, c->M, c->micro_runs, c->acc.p, c->span_cum.p, c->fin.p,
                                  c->fin.p + 7 * S, c->stream));
    c->launches += 2;
    if (mean_out) {
        PZ_CUDA(cudaMemcpyAsync(mean_out, c->fin.p, 7 * S * 8, cudaMemcpyDeviceToHost, c->stream));
        PZ_CUDA(cudaMemcpyAsync(var_out, c->fin.p + 7 * S, 6 * S * 8, cudaMemcpyDeviceToHost, c->stream));
    }
    PZ_CUDA(cudaStreamSynchronize(c->stream));
    return PZ_OK;
}

int pz_micro_arrays(pz_ctx *c, double t_lo, double t_hi, double norm, double *out)
{
    if (!c || !out || !(norm > 0.0)) return fail(PZ_ERR_ARG, "pz_micro_arrays: bad arguments");
    int rc = pz_micro_finalize(c, nullptr, nullptr);        // mean / var stay on the device
    if (rc) return rc;
    const size_t S = (size_t)c->M + 1;
    PZ_CUDA(c->arrays.ensure(19 * S));
    PZ_CUDA(launch_micro_arrays(c->M, c->micro_runs, t_lo, t_hi, norm, c->fin.p, c->fin.p + 7 * S,
                                c->arrays.p, c->stream));
    c->launches += 1;
    PZ_CUDA(cudaMemcpyAsync(out, c->arrays.p, 19 * S * 8, cudaMemcpyDeviceToHost, c->stream));
    PZ_CUDA(cudaStreamSynchronize(c->stream));
    return PZ_OK;
}

int pz_set_ps(pz_ctx *c, int32_t M, int32_t num_p, const double *ps, double *pmf_out)
{
    if (!c) return fail(PZ_ERR_ARG, "null context");
    if (M < 0) return fail(PZ_ERR_ARG, "pz_set_ps: M must be >= 0");
    if (num_p < 0 || (num_p > 0 && !ps)) return fail(PZ_ERR_ARG, "pz_set_ps: bad arguments");
    for (int i = 0; i < num_p; ++i)
        if (!(ps[i] >= 0.0 && ps[i] <= 1.0)) return fail(PZ_ERR_ARG, "pz_set_ps: p must lie in [0, 1]");
    PZ_CUDA(cudaSetDevice(c->device));
    const bool want_sf = c->N > 0 && M == c->M;      // only the fused path needs the table
    // same table as last time (a study calls this once per batch of seeds): keep it
    const bool cached = num_p > 0 && c->num_p == num_p && c->pmf_M == M &&
                        std::equal(ps, ps + num_p, c->ps.begin()) &&
                        (!want_sf || c->sf_M == M) && c->pmf.p != nullptr;
    c->num_p = num_p;
    c->pmf_M = M;
    c->ps.assign(ps, ps + num_p);
    c->canon_count = 0; c->canon_last_R = 0;
    if (num_p == 0) return PZ_OK;
    const size_t S = (size_t)M + 1;
    if (cached) {
        if (pmf_out) {
            for (int i = 0; i < num_p; ++i)
                PZ_CUDA(cudaMemcpyAsync(pmf_out + (size_t)c->porder[i] * S, c->pmf.p + (size_t)i * S, S * 8,
                                        cudaMemcpyDeviceToHost, c->stream));
            PZ_CUDA(cudaStreamSynchronize(c->stream));
        }
        return PZ_OK;
    }
    c->porder.resize(num_p);
    for (int i = 0; i < num_p; ++i) c->porder[i] = i;
    std::stable_sort(c->porder.begin(), c->porder.end(),
                     [&](int a, int b) { return ps[a] < ps[b]; });
    std::vector<double> sorted(num_p);
    for (int i = 0; i < num_p; ++i) sorted[i] = ps[c->porder[i]];
    PZ_CUDA(c->ps_dev.ensure(num_p));
    PZ_CUDA(c->pmf.ensure((size_t)num_p * S));
    PZ_CUDA(c->band_lo.ensure(num_p));
    PZ_CUDA(c->band_hi.ensure(num_p));
    PZ_CUDA(c->tband_lo.ensure(num_p));
    PZ_CUDA(c->tband_hi.ensure(num_p));
    PZ_CUDA(c->canon_flags.ensure(num_p));
    if (want_sf) PZ_CUDA(c->sf.ensure((size_t)num_p * S));
    PZ_CUDA(c->porder_dev.ensure(num_p));
    PZ_CUDA(cudaMemcpyAsync(c->ps_dev.p, sorted.data(), (size_t)num_p * 8, cudaMemcpyHostToDevice, c->stream));
    PZ_CUDA(cudaMemcpyAsync(c->porder_dev.p, c->porder.data(), (size_t)num_p * 4, cudaMemcpyHostToDevice, c->stream));
    PZ_CUDA(launch_binomial_pmf(M, num_p, c->ps_dev.p, c->pmf.p, c->band_lo.p, c->band_hi.p,
                                c->tband_lo.p, c->tband_hi.p, want_sf ? c->sf.p : nullptr, c->stream));
    c->launches += want_sf ? 3 : 2;
    c->sf_M = want_sf ? M : -1;
    if (pmf_out)
        for (int i = 0; i < num_p; ++i)
            PZ_CUDA(cudaMemcpyAsync(pmf_out + (size_t)c->porder[i] * S, c->pmf.p + (size_t)i * S, S * 8,
                                    cudaMemcpyDeviceToHost, c->stream));
    PZ_CUDA(cudaStreamSynchronize(c->stream));
    return PZ_OK;
}

int pz_convolve(pz_ctx *c, int32_t num_cols, const double *cols, double *out)
{
    if (!c || num_cols < 0 || (num_cols > 0 && (!cols || !out)))
        return fail(PZ_ERR_ARG, "pz_convolve: bad arguments");
    if (c->num_p == 0) return fail(PZ_ERR_STATE, "pz_convolve: call pz_set_ps first");
    if (num_cols == 0) return PZ_OK;
    PZ_CUDA(cudaSetDevice(c->device));
    const size_t S = (size_t)c->pmf_M + 1;
    const int P = c->num_p;
    PZ_CUDA(c->cols.ensure((size_t)num_cols * S));
    PZ_CUDA(c->cols_out.ensure((size_t)num_cols * P));
    PZ_CUDA(cudaMemcpyAsync(c->cols.p, cols, (size_t)num_cols * S * 8, cudaMemcpyHostToDevice, c->stream));
    PZ_CUDA(launch_convolve(c->pmf_M, P, c->pmf.p, c->band_lo.p, c->band_hi.p, num_cols, c->cols.p,
                            c->cols_out.p, c->stream));
    c->launches += 1;
    std::vector<double> tmp((size_t)num_cols * P);
    PZ_CUDA(cudaMemcpyAsync(tmp.data(), c->cols_out.p, tmp.size() * 8, cudaMemcpyDeviceToHost, c->stream));
    PZ_CUDA(cudaStreamSynchronize(c->stream));
    for (int col = 0; col < num_cols; ++col)
        for (int i = 0; i < P; ++i)
            out[(size_t)col * P + c->porder[i]] = tmp[(size_t)col * P + i];
    return PZ_OK;
}

int pz_canonical_statistics_rows(pz_ctx *c, int32_t M, int spanning, const void *rows,
                                 const double *f, double *out)
{
    if (!c || M < 0 || !rows || !f || !out)
        return fail(PZ_ERR_ARG, "pz_canonical_statistics_rows: bad arguments");
    PZ_CUDA(cudaSetDevice(c->device));
    const size_t S = (size_t)M + 1, rb = spanning ? 53 : 52;
    PZ_CUDA(c->rows.ensure(S * rb));
    PZ_CUDA(c->cols.ensure(S + 8));
    PZ_CUDA(cudaMemcpyAsync(c->rows.p, rows, S * rb, cudaMemcpyHostToDevice, c->stream));
    PZ_CUDA(cudaMemcpyAsync(c->cols.p, f, S * 8, cudaMemcpyHostToDevice, c->stream));
    PZ_CUDA(launch_canon_rows(M, spanning ? 1 : 0, c->rows.p, c->cols.p, c->cols.p + S, c->stream));
    c->launches += 1;
    PZ_CUDA(cudaMemcpyAsync(out, c->cols.p + S, 7 * 8, cudaMemcpyDeviceToHost, c->stream));
    PZ_CUDA(cudaStreamSynchronize(c->stream));
    return PZ_OK;
}

int pz_canon_export(pz_ctx *c, int64_t *count_out, double *mean_out, double *m2_out)
{
    if (!c || !count_out || !mean_out || !m2_out) return fail(PZ_ERR_ARG, "pz_canon_export: bad arguments");
    if (c->num_p == 0) return fail(PZ_ERR_STATE, "pz_canon_export: call pz_set_ps first");
    const size_t cols = (size_t)c->num_p * PZ_CANON_COLS;
    *count_out = c->canon_count;
    for (size_t i = 0; i < cols; ++i) {
        mean_out[i] = c->canon_count ? c->canon_mean[i] : 0.0;
        m2_out[i] = c->canon_count ? c->canon_m2[i] : 0.0;
    }
    return PZ_OK;
}

int pz_canon_merge(pz_ctx *c, int64_t count, const double *mean, const double *m2)
{
    if (!c || count < 0 || !mean || !m2) return fail(PZ_ERR_ARG, "pz_canon_merge: bad arguments");
    if (c->num_p == 0) return fail(PZ_ERR_STATE, "pz_canon_merge: call pz_set_ps first");
    const size_t cols = (size_t)c->num_p * PZ_CANON_COLS;
    c->canon_mean.resize(cols); c->canon_m2.resize(cols);
    chan_merge(c->canon_count, c->canon_mean, c->canon_m2, count, mean, m2);
    return PZ_OK;
}

int pz_canon_reset(pz_ctx *c)
{
    if (!c) return fail(PZ_ERR_ARG, "null context");
    c->canon_count = 0;
    return PZ_OK;
}

int32_t pz_canon_last_count(const pz_ctx *c) { return c ? c->canon_last_R : 0; }

int pz_canon_last_runs(pz_ctx *c, double *out)
{
    if (!c || !out) return fail(PZ_ERR_ARG, "pz_canon_last_runs: bad arguments");
    if (c->canon_last_R == 0) return fail(PZ_ERR_STATE, "pz_canon_last_runs: no fused canonical batch yet");
    PZ_CUDA(cudaSetDevice(c->device));
    PZ_CUDA(cudaMemcpyAsync(out, c->canon_last_ptr,
                            (size_t)c->canon_last_R * c->num_p * PZ_CANON_COLS * 8,
                            cudaMemcpyDeviceToHost, c->stream));
    PZ_CUDA(cudaStreamSynchronize(c->stream));
    return PZ_OK;
}

}  // extern "C"

extern "C" {

int pz_profile(pz_ctx *c, int enable)
{
    if (!c) return fail(PZ_ERR_ARG, "null context");
    c->profiling = enable != 0;
    for (int i = 0; i < PZ_PHASES; ++i) { c->phase_ms[i] = 0; c->phase_launches[i] = 0; }
    return PZ_OK;
}

int pz_profile_read(pz_ctx *c, double *ms_out, int64_t *launches_out)
{
    if (!c || !ms_out || !launches_out) return fail(PZ_ERR_ARG, "pz_profile_read: bad arguments");
    for (int i = 0; i < PZ_PHASES; ++i) { ms_out[i] = c->phase_ms[i]; launches_out[i] = c->phase_launches[i]; }
    return PZ_OK;
}

}  // extern "C"

extern "C" {

int pz_timer_start(pz_ctx *c)
{
    if (!c) return fail(PZ_ERR_ARG, "null context");
    PZ_CUDA(cudaSetDevice(c->device));
    if (!c->timer_a) { PZ_CUDA(cudaEventCreate(&c->timer_a)); PZ_CUDA(cudaEventCreate(&c->timer_b)); }
    PZ_CUDA(cudaEventRecord(c->timer_a, c->stream));
    return PZ_OK;
}

int pz_timer_stop(pz_ctx *c, double *ms_out)
{
    if (!c || !ms_out) return fail(PZ_ERR_ARG, "pz_timer_stop: bad arguments");
    if (!c->timer_a) return fail(PZ_ERR_STATE, "pz_timer_stop: timer not started");
    PZ_CUDA(cudaSetDevice(c->device));
    PZ_CUDA(cudaEventRecord(c->timer_b, c->stream));
    PZ_CUDA(cudaEventSynchronize(c->timer_b));
    float ms = 0.f;
    PZ_CUDA(cudaEventElapsedTime(&ms, c->timer_a, c->timer_b));
    *ms_out = ms;
    return PZ_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------
// cross-GPU exchange: one process per GPU, NCCL over NVLink / NVSwitch.
// The reference reduces the per-task results of a study with bond_reduce over pickles on a
// shared file system (percolate/share/jugfile.py:126-135, 240-244); here every rank folds its
// runs into its context and ONE collective step combines the ranks.  NCCL is loaded at run time
// (dlopen): single-GPU users do not need it, and inside a PyTorch process the copy torch has
// already loaded is the one that is found.
// ---------------------------------------------------------------------------
#include <dlfcn.h>
#include <nccl.h>
#include <mutex>

namespace {
struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                              cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    std::string why;
};
NcclApi g_nccl;
std::once_flag g_nccl_once;

void load_nccl()
{
    const char *names[] = {getenv("PZ_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
        if (!nm || !*nm) continue;
        g_nccl.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.lib) break;
        g_nccl.why = dlerror();
    }
    if (!g_nccl.lib) return;
#define PZ_SYM(field, name)                                                        \
    g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(g_nccl.lib, name)); \
    if (!g_nccl.field) { g_nccl.why = std::string("missing symbol ") + name; g_nccl.lib = nullptr; return; }
    PZ_SYM(GetUniqueId, "ncclGetUniqueId")
    PZ_SYM(CommInitRank, "ncclCommInitRank")
    PZ_SYM(CommDestroy, "ncclCommDestroy")
    PZ_SYM(AllReduce, "ncclAllReduce")
    PZ_SYM(AllGather, "ncclAllGather")
    PZ_SYM(GroupStart, "ncclGroupStart")
    PZ_SYM(GroupEnd, "ncclGroupEnd")
    PZ_SYM(GetErrorString, "ncclGetErrorString")
#undef PZ_SYM
}

int need_nccl()
{
    std::call_once(g_nccl_once, load_nccl);
    if (!g_nccl.lib) return fail(PZ_ERR_STATE, "NCCL could not be loaded: " + g_nccl.why);
    return PZ_OK;
}
}  // namespace

#define PZ_NCCL(expr)                                                                     \
    do {                                                                                  \
        ncclResult_t _r = (expr);                                                         \
        if (_r != ncclSuccess)                                                            \
            return fail(PZ_ERR_CUDA, std::string(#expr) + ": " + g_nccl.GetErrorString(_r)); \
    } while (0)

extern "C" {

int pz_comm_unique_id(void *id_out)
{
    if (!id_out) return fail(PZ_ERR_ARG, "pz_comm_unique_id: id_out is NULL");
    int rc = need_nccl(); if (rc) return rc;
    static_assert(sizeof(ncclUniqueId) == PZ_COMM_ID_BYTES, "PZ_COMM_ID_BYTES");
    ncclUniqueId id;
    PZ_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(id_out, &id, sizeof(id));
    return PZ_OK;
}

int pz_comm_init(pz_ctx *c, int world, int rank, const void *id)
{
    if (!c || !id || world < 1 || rank < 0 || rank >= world)
        return fail(PZ_ERR_ARG, "pz_comm_init: bad arguments");
    int rc = need_nccl(); if (rc) return rc;
    PZ_CUDA(cudaSetDevice(c->device));
    pz_comm_destroy(c);
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    ncclComm_t comm = nullptr;
    PZ_NCCL(g_nccl.CommInitRank(&comm, world, uid, rank));
    c->comm = comm;
    c->comm_world = world;
    c->comm_rank = rank;
    return PZ_OK;
}

int pz_comm_destroy(pz_ctx *c)
{
    if (!c) return fail(PZ_ERR_ARG, "null context");
    if (c->comm && g_nccl.lib) {
        cudaSetDevice(c->device);
        cudaStreamSynchronize(c->stream);
        g_nccl.CommDestroy((ncclComm_t)c->comm);
    }
    c->comm = nullptr;
    c->comm_world = 1;
    c->comm_rank = 0;
    return PZ_OK;
}

int pz_comm_world(const pz_ctx *c) { return c && c->comm ? c->comm_world : 1; }
int pz_comm_rank(const pz_ctx *c) { return c && c->comm ? c->comm_rank : 0; }

int pz_allreduce(pz_ctx *c)
{
    if (!c) return fail(PZ_ERR_ARG, "null context");
    if (!c->comm) return fail(PZ_ERR_STATE, "pz_allreduce: call pz_comm_init first");
    if (c->N == 0) return fail(PZ_ERR_STATE, "no graph set");
    PZ_CUDA(cudaSetDevice(c->device));
    ncclComm_t comm = (ncclComm_t)c->comm;
    const int W = c->comm_world;
    int rc = ensure_acc(c); if (rc) return rc;
    const size_t words = ((size_t)c->M + 1) * PZ_ACC_WORDS;
    const size_t cols = (size_t)c->num_p * PZ_CANON_COLS;
    const size_t pay = 1 + 2 * cols;                        // count, mean[cols], M2[cols]
    const size_t host_need = 1 + pay + (size_t)W * pay;     // run count word + send + recv
    if (c->comm_host_cap < host_need) {
        if (c->comm_host) cudaFreeHost(c->comm_host);
        c->comm_host = nullptr; c->comm_host_cap = 0;
        PZ_CUDA(cudaMallocHost(&c->comm_host, host_need * 8));
        c->comm_host_cap = host_need;
    }
    // the run count rides in the word behind the accumulator block: one integer all-reduce
    int64_t *runs_word = reinterpret_cast<int64_t *>(c->comm_host);
    *runs_word = c->micro_runs;
    PZ_CUDA(cudaMemcpyAsync(c->acc.p + words, runs_word, 8, cudaMemcpyHostToDevice, c->stream));
    double *send_h = c->comm_host + 1, *recv_h = c->comm_host + 1 + pay;
    if (cols) {
        send_h[0] = (double)c->canon_count;
        for (size_t i = 0; i < cols; ++i) {
            send_h[1 + i] = c->canon_count ? c->canon_mean[i] : 0.0;
            send_h[1 + cols + i] = c->canon_count ? c->canon_m2[i] : 0.0;
        }
        PZ_CUDA(c->comm_send.ensure(pay));
        PZ_CUDA(c->comm_recv.ensure((size_t)W * pay));
        PZ_CUDA(cudaMemcpyAsync(c->comm_send.p, send_h, pay * 8, cudaMemcpyHostToDevice, c->stream));
    }
    PZ_NCCL(g_nccl.GroupStart());
    // exact: every accumulator word holds a 32-bit limb, so the word-wise sum cannot carry
    PZ_NCCL(g_nccl.AllReduce(c->acc.p, c->acc.p, words + 1, ncclInt64, ncclSum, comm, c->stream));
    if (cols)
        PZ_NCCL(g_nccl.AllGather(c->comm_send.p, c->comm_recv.p, pay, ncclFloat64, comm, c->stream));
    PZ_NCCL(g_nccl.GroupEnd());
    PZ_CUDA(cudaMemcpyAsync(runs_word, c->acc.p + words, 8, cudaMemcpyDeviceToHost, c->stream));
    if (cols)
        PZ_CUDA(cudaMemcpyAsync(recv_h, c->comm_recv.p, (size_t)W * pay * 8, cudaMemcpyDeviceToHost, c->stream));
    PZ_CUDA(cudaStreamSynchronize(c->stream));
    c->launches += cols ? 2 : 1;
    c->micro_runs = *runs_word;
    if (cols) {
        // rank-ordered Chan merge (the arithmetic of bond_reduce): bit-identical on every rank
        c->canon_count = 0;
        c->canon_mean.assign(cols, 0.0);
        c->canon_m2.assign(cols, 0.0);
        for (int r = 0; r < W; ++r) {
            const double *p = recv_h + (size_t)r * pay;
            chan_merge(c->canon_count, c->canon_mean, c->canon_m2, (int64_t)llround(p[0]), p + 1, p + 1 + cols);
        }
    }
    return PZ_OK;
}

}  // extern "C"
