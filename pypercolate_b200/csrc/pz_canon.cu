// pz_canon.cu -- canonical ensemble: binomial weights and the contractions
// with them (fp64, CUDA cores; sm_100a).
//
//   binomial_pmf      _binomial_pmf, percolate/percolate.py:1067-1109
//   convolve          canonical_averages, percolate/percolate.py:1196-1221
//   canon_rows        bond_canonical_statistics on host rows, percolate/hpc.py:488-515
//   canon_runs        the same contraction for every run of a batch straight
//                     from the merge records (nothing per-run materialised);
//                     tight band with a verified error bound + exact fallback
//   canon_reduce      bond_initialize_canonical_averages + bond_reduce over the
//                     runs of a batch, percolate/hpc.py:607-635, 664-702
#include "pz_common.cuh"
#include "pz_internal.h"

namespace pz {

// ---------------------------------------------------------------------------
// _binomial_pmf: one thread per (p, direction).  The recurrence is evaluated
// in the reference's operation order with explicitly rounded multiplies and
// divides (no contraction), so every un-normalised weight is bit-identical to
// the reference's; only the normalising sum is associated differently.
// ---------------------------------------------------------------------------
__global__ void binomial_pmf_kernel(int32_t M, int32_t P, const double *ps, double *pmf)
{
    const int pi = blockIdx.x * blockDim.x + threadIdx.x;
    if (pi >= P) return;
    const double p = ps[pi];
    const double n = (double)M;
    const long long nmax = (long long)rint(p * n);          // np.round: half to even
    double *ret = pmf + (size_t)pi * ((size_t)M + 1);
    if (blockIdx.y == 0) {
        double v = 1.0;
        ret[nmax] = v;
        const double q = 1.0 - p;
        for (long long i = nmax + 1; i <= M; ++i) {
            // ret[i] = ret[i-1] * (n - i + 1.0) / i * p / (1.0 - p)
            v = __ddiv_rn(__dmul_rn(__ddiv_rn(__dmul_rn(v, n - (double)i + 1.0), (double)i), p), q);
            ret[i] = v;
        }
    } else {
        double v = 1.0;
        const double q = 1.0 - p;
        for (long long i = nmax - 1; i >= 0; --i) {
            // ret[i] = ret[i+1] * (i + 1.0) / (n - i) * (1.0 - p) / p
            v = __ddiv_rn(__dmul_rn(__ddiv_rn(__dmul_rn(v, (double)i + 1.0), (double)(M - i)), q), p);
            ret[i] = v;
        }
    }
}

__device__ __forceinline__ double block_sum(double v, double *sh)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = 16; k > 0; k >>= 1) v += __shfl_xor_sync(0xffffffffu, v, k);
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
    return t;
}

// normalise (ret / ret.sum()) and find two bands per p:
//   exact band  [xlo, xhi]  outside which the weight is exactly zero (underflowed):
//                           banded sums over it drop no term at all;
//   tight band  [tlo, thi]  outside which the weight is below PMF_TIGHT: the dropped mass
//                           is at most (M+1) * PMF_TIGHT, which bounds the error of a banded
//                           sum absolutely (see canon_runs_kernel).
static constexpr double PMF_TIGHT = 1e-40;
__global__ void __launch_bounds__(256) pmf_normalize_kernel(int32_t M, double *pmf, int32_t *xlo,
                                                             int32_t *xhi, int32_t *tlo, int32_t *thi)
{
    __shared__ double sh[8];
    __shared__ int s_lo, s_hi, s_tlo, s_thi;
    double *ret = pmf + (size_t)blockIdx.x * ((size_t)M + 1);
    // chunked partial sums keep the association close to a pairwise sum
    double part = 0.0;
    for (int i = threadIdx.x; i <= M; i += blockDim.x) part += ret[i];
    const double s = block_sum(part, sh);
    if (threadIdx.x == 0) { s_lo = M; s_hi = 0; s_tlo = M; s_thi = 0; }
    __syncthreads();
    int lo = M, hi = 0, lo2 = M, hi2 = 0;
    for (int i = threadIdx.x; i <= M; i += blockDim.x) {
        const double v = __ddiv_rn(ret[i], s);
        ret[i] = v;
        if (v > 0.0) { lo = min(lo, i); hi = max(hi, i); }
        if (v >= PMF_TIGHT) { lo2 = min(lo2, i); hi2 = max(hi2, i); }
    }
    atomicMin(&s_lo, lo);
    atomicMax(&s_hi, hi);
    atomicMin(&s_tlo, lo2);
    atomicMax(&s_thi, hi2);
    __syncthreads();
    if (threadIdx.x == 0) {
        xlo[blockIdx.x] = min(s_lo, s_hi);
        xhi[blockIdx.x] = max(s_lo, s_hi);
        tlo[blockIdx.x] = min(s_tlo, s_thi);
        thi[blockIdx.x] = max(s_tlo, s_thi);
    }
}

// survival function sf[n] = sum_{m >= n} pmf[m] (one CTA per p, accumulated from
// the top so that small tails keep their relative accuracy): the canonical
// percolation probability of a run is sf[first spanning n] (hpc.py:495-499)
static constexpr int SF_ITEMS = 8;
__global__ void __launch_bounds__(256) pmf_sf_kernel(int32_t M, const double *pmf, double *sf)
{
    __shared__ double wtot[8];
    __shared__ double carry;
    const double *f = pmf + (size_t)blockIdx.x * ((size_t)M + 1);
    double *out = sf + (size_t)blockIdx.x * ((size_t)M + 1);
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    if (t == 0) carry = 0.0;
    __syncthreads();
    for (long long top = M; top >= 0; top -= 256 * SF_ITEMS) {
        const long long first = top - (long long)t * SF_ITEMS;       // this thread: first, first-1, ...
        double v[SF_ITEMS], mine = 0.0;
#pragma unroll
        for (int i = 0; i < SF_ITEMS; ++i) {
            const long long n = first - i;
            v[i] = n >= 0 ? f[n] : 0.0;
            mine += v[i];
        }
        double incl = mine;
        for (int k = 1; k < 32; k <<= 1) {
            const double o = __shfl_up_sync(0xffffffffu, incl, k);
            if (lane >= k) incl += o;
        }
        if (lane == 31) wtot[warp] = incl;
        __syncthreads();
        double run = carry + (incl - mine);
        for (int w = 0; w < warp; ++w) run += wtot[w];
#pragma unroll
        for (int i = 0; i < SF_ITEMS; ++i) {
            const long long n = first - i;
            run += v[i];
            if (n >= 0) out[n] = run;
        }
        __syncthreads();
        if (t == 255) carry = run;
        __syncthreads();
    }
}

cudaError_t launch_binomial_pmf(int32_t M, int32_t P, const double *ps_dev, double *pmf,
                                int32_t *xlo, int32_t *xhi, int32_t *tlo, int32_t *thi, double *sf,
                                cudaStream_t s)
{
    if (P <= 0) return cudaSuccess;
    dim3 grid((P + 31) / 32, 2);
    binomial_pmf_kernel<<<grid, 32, 0, s>>>(M, P, ps_dev, pmf);
    pmf_normalize_kernel<<<P, 256, 0, s>>>(M, pmf, xlo, xhi, tlo, thi);
    if (sf) pmf_sf_kernel<<<P, 256, 0, s>>>(M, pmf, sf);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// out[c][p] = sum_n pmf[p][n] * cols[c][n]   (one CTA per (p, c))
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) convolve_kernel(int32_t M, int32_t P, const double *pmf,
                                                        const int32_t *band_lo, const int32_t *band_hi,
                                                        const double *cols, double *out)
{
    __shared__ double sh[8];
    const int p = blockIdx.x, c = blockIdx.y;
    const double *f = pmf + (size_t)p * ((size_t)M + 1);
    const double *x = cols + (size_t)c * ((size_t)M + 1);
    double part = 0.0;
    for (int i = band_lo[p] + threadIdx.x; i <= band_hi[p]; i += blockDim.x)
        part += __dmul_rn(f[i], x[i]);
    const double s = block_sum(part, sh);
    if (threadIdx.x == 0) out[(size_t)c * P + p] = s;
}

cudaError_t launch_convolve(int32_t M, int32_t P, const double *pmf, const int32_t *band_lo,
                            const int32_t *band_hi, int32_t num_cols, const double *cols,
                            double *out, cudaStream_t s)
{
    if (P <= 0 || num_cols <= 0) return cudaSuccess;
    dim3 grid(P, num_cols);
    convolve_kernel<<<grid, 256, 0, s>>>(M, P, pmf, band_lo, band_hi, cols, out);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// bond_canonical_statistics on packed host rows of ONE run (hpc.py:488-515)
// ---------------------------------------------------------------------------
template <bool SPANNING>
__global__ void __launch_bounds__(256) canon_rows_kernel(int32_t M, const uint8_t *rows,
                                                          const double *f, double *out)
{
    constexpr int RB = SPANNING ? 53 : 52;
    constexpr int MOFF = SPANNING ? 9 : 8;
    __shared__ double sh[8];
    double acc[7];
#pragma unroll
    for (int q = 0; q < 7; ++q) acc[q] = 0.0;
    for (int i = threadIdx.x; i <= M; i += blockDim.x) {
        const uint8_t *r = rows + (size_t)i * RB;
        const double w = f[i];
        if (SPANNING) acc[0] += __dmul_rn(w, (double)r[8]);
        uint32_t mx = 0;
        for (int b = 0; b < 4; ++b) mx |= (uint32_t)r[MOFF + b] << (8 * b);
        acc[1] += __dmul_rn(w, (double)mx);
        for (int k = 0; k < 5; ++k) {
            uint64_t m = 0;
            for (int b = 0; b < 8; ++b) m |= (uint64_t)r[MOFF + 4 + 8 * k + b] << (8 * b);
            acc[2 + k] += __dmul_rn(w, (double)m);
        }
    }
    for (int q = 0; q < 7; ++q) {
        const double s = block_sum(acc[q], sh);
        if (threadIdx.x == 0) out[q] = s;
    }
}

cudaError_t launch_canon_rows(int32_t M, int spanning, const uint8_t *rows, const double *f,
                              double *out, cudaStream_t s)
{
    if (spanning) canon_rows_kernel<true><<<1, 256, 0, s>>>(M, rows, f, out);
    else canon_rows_kernel<false><<<1, 256, 0, s>>>(M, rows, f, out);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// canon_runs: out[run][p][7] = sum_n pmf_p[n] * Q_run[n] for all runs of a
// batch, straight from the merge records.
//
// One lane owns one run (a warp = 32 runs in lock step), one warp owns a chunk
// of PC probabilities whose weights live in 6*PC register accumulators.  The
// warp starts at the last checkpoint (run state every `ckpt_every` rows, left
// by checkpoint_kernel) before the chunk's band and walks to its end; records
// and weights of a tile of rows are staged in shared memory.
//
// Column 0 (percolation probability) is not summed: the spanning flag is a
// step function, so its contraction is the survival function at the first
// spanning n -- one table look-up.
//
// Two passes.  Pass 0 sums over the TIGHT band (weights >= 1e-40); the dropped
// terms are bounded by err = (M+1) * 1e-40 * 2^64 in absolute value, so every
// result >= err * 1e11 is within 1e-11 relative of the full sum.  A chunk with
// a smaller result raises a flag; pass 1 (same kernel, exact band = every
// non-zero weight) recomputes only flagged chunks, so no term is ever dropped
// from a result that could notice it.
// ---------------------------------------------------------------------------
// probabilities per warp: 8 (48 fp64 accumulators) when the batch has enough runs to fill the
// GPU, 4 for short batches of large graphs (twice the warps; c4: 1e3 runs of L = 1024)


template <class RecT, int PC>
__global__ void __launch_bounds__(128, 3) canon_runs_kernel(StatsArgs a, int32_t P, int32_t nchunks,
                                                             const double *pmf, const double *sf,
                                                             const int32_t *xlo, const int32_t *xhi,
                                                             const int32_t *tlo, const int32_t *thi,
                                                             const int32_t *porder,
                                                             const RunState *ckpt, int ckpt_every,
                                                             int n_ckpt, double *out, int *flags,
                                                             int pass)
{
    constexpr int TILE = sizeof(RecT) == 4 ? 32 : 16;
    __shared__ RecT tile_all[4][32][TILE + 1];
    __shared__ double ftile_all[4][PC][TILE];         // weights of the rows of the current tile
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    RecT (*tile)[TILE + 1] = tile_all[warp];
    double (*ftile)[TILE] = ftile_all[warp];
    const long long wid = (long long)blockIdx.x * 4 + warp;
    const int rg = (int)(wid / nchunks), ch = (int)(wid % nchunks);
    const int run0 = rg * 32;
    if (run0 >= a.R) return;
    if (pass == 1 && flags[ch] == 0) return;          // nothing in this chunk needs the exact band
    const int run = run0 + lane;
    const bool live = run < a.R;
    const int M = a.M;
    const size_t S = (size_t)M + 1;
    const RecT *recs = reinterpret_cast<const RecT *>(a.recs);

    // chunk band (sorted-p order)
    const int p0 = ch * PC, np = min(PC, P - p0);
    const int32_t *blo = pass == 0 ? tlo : xlo, *bhi = pass == 0 ? thi : xhi;
    int lo = M, hi = 0;
    bool narrowed = false;
#pragma unroll
    for (int k = 0; k < PC; ++k)
        if (k < np) {
            lo = min(lo, blo[p0 + k]); hi = max(hi, bhi[p0 + k]);
            narrowed |= (tlo[p0 + k] != xlo[p0 + k]) || (thi[p0 + k] != xhi[p0 + k]);
        }
    const int ck = lo / ckpt_every;                 // checkpoint row = ck * ckpt_every <= lo
    const int row_start = ck * ckpt_every;          // state is the one AFTER this row
    RunState st;
    st.init((uint32_t)a.N);
    if (live) st = ckpt[(size_t)run * n_ckpt + ck];
    const uint32_t nspan = (live && a.spanning) ? a.nspan[run] : NSPAN_NEVER;

    double acc[PC][6];
#pragma unroll
    for (int k = 0; k < PC; ++k)
#pragma unroll
        for (int q = 0; q < 6; ++q) acc[k][q] = 0.0;

    double x[6];
    bool dirty = true;
    auto refresh = [&]() {
        if (dirty) {
            uint64_t m[5];
            st.moments((uint32_t)a.N, m);
            x[0] = (double)st.mx;
#pragma unroll
            for (int k = 0; k < 5; ++k) x[1 + k] = (double)m[k];
            dirty = false;
        }
    };

    // rows row_start .. hi; the state after row_start comes from the checkpoint,
    // row n >= 1 applies record n-1.  Inside the union band of the chunk every
    // weight is used (outside its own exact band a weight is exactly 0.0).
    int row = row_start;
    if (row >= lo) {
        refresh();
#pragma unroll
        for (int k = 0; k < PC; ++k) {
            const double f = __ldg(&pmf[(size_t)min(p0 + k, P - 1) * S + row]);
#pragma unroll
            for (int q = 0; q < 6; ++q) acc[k][q] = fma(f, x[q], acc[k][q]);
        }
    }
    while (row < hi) {
        __syncwarp();
        if (lane < TILE) {
            const int idx = row + lane;           // record idx belongs to row idx+1
            for (int rr = 0; rr < 32; ++rr) {
                RecT v = 0;
                if (run0 + rr < a.R && idx < M) v = __ldg(&recs[(size_t)(run0 + rr) * M + idx]);
                tile[rr][lane] = v;
            }
#pragma unroll
            for (int k = 0; k < PC; ++k)
                ftile[k][lane] = (idx + 1 <= hi && idx + 1 >= lo)
                                     ? __ldg(&pmf[(size_t)min(p0 + k, P - 1) * S + idx + 1]) : 0.0;
        }
        __syncwarp();
        const int cnt = min(TILE, hi - row);
        for (int j = 0; j < cnt; ++j) {
            const RecT r = tile[lane][j];
            const int nrow = row + 1 + j;
            if (RecCodec<RecT>::valid(r)) {
                st.merge(RecCodec<RecT>::w_small(r), RecCodec<RecT>::w_large(r));
                dirty = true;
            }
            if (nrow >= lo) {
                refresh();
#pragma unroll
                for (int k = 0; k < PC; ++k) {
                    const double f = ftile[k][j];
#pragma unroll
                    for (int q = 0; q < 6; ++q) acc[k][q] = fma(f, x[q], acc[k][q]);
                }
            }
        }
        row += cnt;
    }
    // error bound of the tight band: dropped mass * largest possible value
    const double floor_ok = (double)(M + 1) * PMF_TIGHT * 18446744073709551616.0 * 1e11;
    bool small = false;
    if (live) {
        for (int k = 0; k < np; ++k) {
            double *o = out + ((size_t)run * P + porder[p0 + k]) * 7;
            o[0] = (nspan != NSPAN_NEVER && nspan <= (uint32_t)M) ? __ldg(&sf[(size_t)(p0 + k) * S + nspan]) : 0.0;
            for (int q = 0; q < 6; ++q) {
                o[1 + q] = acc[k][q];
                small |= acc[k][q] < floor_ok;
            }
        }
    }
    if (pass == 0 && narrowed && __any_sync(0xffffffffu, small) && lane == 0) atomicOr(&flags[ch], 1);
}

cudaError_t launch_canon_runs(const StatsArgs &a, int32_t P, const double *pmf, const double *sf,
                              const int32_t *xlo, const int32_t *xhi, const int32_t *tlo,
                              const int32_t *thi, const int32_t *porder, const RunState *ckpt,
                              int ckpt_every, int n_ckpt, double *out, int *flags, cudaStream_t s)
{
    if (a.R <= 0 || P <= 0) return cudaSuccess;
    int sms = 148;
    {
        int dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const long long groups = (a.R + 31) / 32;
    const bool narrow = groups * ((P + 7) / 8) < (long long)sms * 12;      // cannot fill 12 warps per SM
    const int pc = narrow ? 4 : 8;
    const int nchunks = (P + pc - 1) / pc;
    const long long warps = groups * nchunks;
    const int grid = (int)((warps + 3) / 4);
    cudaError_t e = cudaMemsetAsync(flags, 0, sizeof(int) * nchunks, s);
    if (e != cudaSuccess) return e;
    for (int pass = 0; pass < 2; ++pass) {
#define PZ_CANON_LAUNCH(T, C)                                                                          \
        canon_runs_kernel<T, C><<<grid, 128, 0, s>>>(a, P, nchunks, pmf, sf, xlo, xhi, tlo, thi, porder, \
                                                     ckpt, ckpt_every, n_ckpt, out, flags, pass)
        if (a.rec64) { if (narrow) PZ_CANON_LAUNCH(uint64_t, 4); else PZ_CANON_LAUNCH(uint64_t, 8); }
        else { if (narrow) PZ_CANON_LAUNCH(uint32_t, 4); else PZ_CANON_LAUNCH(uint32_t, 8); }
#undef PZ_CANON_LAUNCH
    }
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// canon_reduce: (mean, M2) over the R runs of a batch for each of P*7 columns.
// Shifted-data form with the first run as pivot: d = x - x0,
// mean = x0 + sum(d)/R, M2 = sum(d^2) - sum(d)^2/R.  Fixed association ->
// deterministic; identical runs give M2 == 0 exactly, as the reference's
// pairwise merges do (percolate/hpc.py:677-684).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) canon_reduce_kernel(int32_t R, int32_t cols, const double *runs,
                                                            double *mean, double *m2)
{
    __shared__ double sh[8];
    const int c = blockIdx.x;
    const double x0 = runs[c];
    double p1 = 0.0, p2 = 0.0;
    for (int r = threadIdx.x; r < R; r += blockDim.x) {
        const double d = runs[(size_t)r * cols + c] - x0;
        p1 += d;
        p2 += d * d;
    }
    const double s1 = block_sum(p1, sh);
    const double s2 = block_sum(p2, sh);
    if (threadIdx.x == 0) {
        mean[c] = x0 + s1 / (double)R;
        const double v = s2 - s1 * s1 / (double)R;
        m2[c] = v > 0.0 ? v : 0.0;
    }
}

cudaError_t launch_canon_reduce(int32_t R, int32_t cols, const double *runs, double *mean,
                                double *m2, cudaStream_t s)
{
    if (R <= 0 || cols <= 0) return cudaSuccess;
    canon_reduce_kernel<<<cols, 256, 0, s>>>(R, cols, runs, mean, m2);
    return cudaGetLastError();
}

}  // namespace pz
