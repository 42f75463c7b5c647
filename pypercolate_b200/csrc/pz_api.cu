// pz_api.cu -- the C-ABI of include/pz.h: context, scratch management and the
// orchestration of the kernels in pz_sweep.cu / pz_stats.cu / pz_rng.cu /
// pz_canon.cu.  No torch types; everything runs on the context's own stream.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>

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
    size_t chunk_bytes = (size_t)4 << 30;

    // graph
    int32_t N = 0, M = 0;
    bool spanning = false;
    int any3 = 0;
    DevBuf<uint32_t> edges32;
    DevBuf<uint2> edges64;
    DevBuf<uint32_t> sides2;

    // per-chunk scratch
    DevBuf<int32_t> perms;
    DevBuf<unsigned char> recs;
    DevBuf<uint32_t> nspan;
    DevBuf<uint32_t> gscratch;
    DevBuf<uint8_t> rows;
    DevBuf<uint32_t> seeds;

    // micro accumulators
    DevBuf<unsigned long long> acc;       // (M+1) * PZ_ACC_WORDS
    DevBuf<unsigned long long> span_hist; // M + 2
    int64_t micro_runs = 0;

    // canonical
    int32_t num_p = 0;
    std::vector<double> ps;
    DevBuf<double> pmf;                   // [num_p][M+1]
    std::vector<int32_t> band_lo, band_hi;
    DevBuf<double> dense;                 // scratch of the contraction
    DevBuf<double> canon_runs;            // [R][num_p][7] of the last fused call
    int32_t canon_last_R = 0;
    int64_t canon_count = 0;
    std::vector<double> canon_mean, canon_m2;

    int64_t launches = 0;
};

// implemented in the other translation units
namespace pz {
cudaError_t launch_accumulate(const StatsArgs &a, unsigned long long *acc,
                              unsigned long long *span_hist, cudaStream_t s, int *launches);
cudaError_t launch_micro_finalize(int32_t N, int32_t M, int64_t runs, const unsigned long long *acc,
                                  const unsigned long long *span_hist, double *mean, double *var,
                                  cudaStream_t s);
cudaError_t launch_binomial_pmf(int32_t M, int32_t num_p, const double *ps_dev, double *pmf,
                                cudaStream_t s);
cudaError_t launch_convolve(int32_t M, int32_t num_p, const double *pmf, int32_t num_cols,
                            const double *cols, double *out, cudaStream_t s);
cudaError_t launch_canon_runs(const StatsArgs &a, int32_t num_p, const double *pmf,
                              const int32_t *band_lo, const int32_t *band_hi, double *out,
                              cudaStream_t s, int *launches);
cudaError_t launch_perm_philox(int32_t M, int32_t R, const uint32_t *seeds, int32_t *perms,
                               cudaStream_t s, int *launches);
cudaError_t launch_perm_mt19937(int32_t M, int32_t R, const uint32_t *seeds, int32_t *perms,
                                cudaStream_t s, int *launches);
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
    if (const char *e = getenv("PZ_FORCE_STORE")) c->force_kind = atoi(e);
    if (const char *e = getenv("PZ_CHUNK_BYTES")) c->chunk_bytes = (size_t)atoll(e);
    *out = c;
    return PZ_OK;
}

void pz_destroy(pz_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    c->edges32.release(); c->edges64.release(); c->sides2.release();
    c->perms.release(); c->recs.release(); c->nspan.release(); c->gscratch.release();
    c->rows.release(); c->seeds.release(); c->acc.release(); c->span_hist.release();
    c->pmf.release(); c->dense.release(); c->canon_runs.release();
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
    PZ_CUDA(cudaStreamSynchronize(c->stream));
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

    std::vector<uint2> e64((size_t)std::max(M, 1));
    for (int32_t e = 0; e < M; ++e) e64[e] = make_uint2((uint32_t)eu[e], (uint32_t)ev[e]);
    PZ_CUDA(c->edges64.ensure(e64.size()));
    PZ_CUDA(cudaMemcpyAsync(c->edges64.p, e64.data(), e64.size() * sizeof(uint2),
                            cudaMemcpyHostToDevice, c->stream));
    std::vector<uint32_t> e32;
    if (N <= 65536) {
        e32.resize((size_t)std::max(M, 1));
        for (int32_t e = 0; e < M; ++e) e32[e] = (uint32_t)eu[e] | ((uint32_t)ev[e] << 16);
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
            s2[x >> 4] |= m << ((x & 15) * 2);
        }
        c->any3 = any3;
        PZ_CUDA(c->sides2.ensure(s2.size()));
        PZ_CUDA(cudaMemcpyAsync(c->sides2.p, s2.data(), s2.size() * 4,
                                cudaMemcpyHostToDevice, c->stream));
    }
    PZ_CUDA(cudaStreamSynchronize(c->stream));
    // graph change invalidates everything derived from it
    c->micro_runs = 0;
    c->acc.release(); c->span_hist.release();
    c->num_p = 0; c->pmf.release();
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

static int sweep_chunk(pz_ctx *c, int32_t R, int perm_mode, const void *perm_src, size_t run0,
                       Chunk *out)
{
    const int32_t M = c->M;
    const int32_t *perms_dev = nullptr;
    const size_t pm = (size_t)R * (size_t)std::max(M, 1);
    if (perm_mode == PZ_PERM_DEVICE) {
        perms_dev = (const int32_t *)perm_src + run0 * (size_t)M;
    } else {
        PZ_CUDA(c->perms.ensure(pm));
        perms_dev = c->perms.p;
        if (perm_mode == PZ_PERM_HOST) {
            if (M > 0)
                PZ_CUDA(cudaMemcpyAsync(c->perms.p, (const int32_t *)perm_src + run0 * (size_t)M,
                                        (size_t)R * M * 4, cudaMemcpyHostToDevice, c->stream));
        } else {
            PZ_CUDA(c->seeds.ensure((size_t)R));
            PZ_CUDA(cudaMemcpyAsync(c->seeds.p, (const uint32_t *)perm_src + run0, (size_t)R * 4,
                                    cudaMemcpyHostToDevice, c->stream));
            int l = 0;
            if (perm_mode == PZ_PERM_PHILOX)
                PZ_CUDA(launch_perm_philox(M, R, c->seeds.p, c->perms.p, c->stream, &l));
            else
                PZ_CUDA(launch_perm_mt19937(M, R, c->seeds.p, c->perms.p, c->stream, &l));
            c->launches += l;
        }
    }
    SweepPlan plan = plan_sweep(c->N, R, c->sms, c->smem_optin, c->force_kind);
    if (plan.kind != STORE_G32 && c->N > 65536)
        return fail(PZ_ERR_ARG, "forced shared-memory store needs N <= 65536");
    const bool rec64 = plan.kind == STORE_G32;
    PZ_CUDA(c->recs.ensure(pm * (rec64 ? 8 : 4)));
    PZ_CUDA(c->nspan.ensure((size_t)R));
    if (plan.gscratch_bytes) PZ_CUDA(c->gscratch.ensure(plan.gscratch_bytes / 4));

    SweepArgs sa{};
    sa.N = c->N; sa.M = M; sa.R = R;
    sa.edges = rec64 ? (const void *)c->edges64.p : (const void *)c->edges32.p;
    sa.sides2 = c->spanning ? c->sides2.p : nullptr;
    sa.any3 = c->any3;
    sa.perms = perms_dev;
    sa.recs = c->recs.p;
    sa.nspan = c->nspan.p;
    sa.gscratch = c->gscratch.p;
    sa.claim_log2 = plan.claim_log2;
    PZ_CUDA(launch_sweep(plan, sa, c->stream));
    c->launches += 1;

    out->plan = plan;
    out->perms_dev = perms_dev;
    out->stats = StatsArgs{c->N, M, R, rec64 ? 1 : 0, c->recs.p, c->nspan.p, perms_dev,
                           c->spanning ? 1 : 0};
    return PZ_OK;
}

static int check_run_args(pz_ctx *c, int32_t R, int perm_mode, const void *perm_src)
{
    if (!c) return fail(PZ_ERR_ARG, "null context");
    if (c->N == 0) return fail(PZ_ERR_STATE, "no graph set (call pz_set_graph first)");
    if (R < 0) return fail(PZ_ERR_ARG, "R must be >= 0");
    if (perm_mode < PZ_PERM_HOST || perm_mode > PZ_PERM_PHILOX)
        return fail(PZ_ERR_ARG, "unknown perm_mode");
    if (R > 0 && !perm_src && c->M > 0) return fail(PZ_ERR_ARG, "perm_src is NULL");
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
    size_t chunk = std::max<size_t>(1, c->chunk_bytes / run_bytes);
    for (size_t r0 = 0; r0 < (size_t)R; r0 += chunk) {
        const int32_t rc_n = (int32_t)std::min(chunk, (size_t)R - r0);
        Chunk ch;
        rc = sweep_chunk(c, rc_n, perm_mode, perm_src, r0, &ch);
        if (rc) return rc;
        PZ_CUDA(c->rows.ensure((size_t)rc_n * run_bytes));
        PZ_CUDA(launch_expand_rows(ch.stats, c->rows.p, c->stream));
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

// ---- TEMPORARY stubs (replaced as the kernels land) -------------------------
extern "C" {
#define PZ_STUB(name, ...) int name(__VA_ARGS__) { return fail(PZ_ERR_STATE, #name ": not implemented yet"); }
PZ_STUB(pz_make_perms, pz_ctx *, int32_t, int, const uint32_t *, int32_t *, int)
PZ_STUB(pz_run_fused, pz_ctx *, int32_t, int, const void *, int)
PZ_STUB(pz_reset_accumulators, pz_ctx *)
int64_t pz_micro_runs(const pz_ctx *c) { return c ? c->micro_runs : 0; }
PZ_STUB(pz_micro_export, pz_ctx *, uint64_t *, int)
PZ_STUB(pz_micro_import, pz_ctx *, const uint64_t *, int, int64_t)
PZ_STUB(pz_micro_finalize, pz_ctx *, double *, double *)
PZ_STUB(pz_set_ps, pz_ctx *, int32_t, const double *, double *)
PZ_STUB(pz_convolve, pz_ctx *, int32_t, const double *, double *)
PZ_STUB(pz_canonical_statistics_rows, pz_ctx *, const void *, const double *, double *)
PZ_STUB(pz_canon_export, pz_ctx *, int64_t *, double *, double *)
PZ_STUB(pz_canon_merge, pz_ctx *, int64_t, const double *, const double *)
PZ_STUB(pz_canon_last_runs, pz_ctx *, double *)
}
