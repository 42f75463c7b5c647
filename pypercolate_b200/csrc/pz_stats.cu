// pz_stats.cu -- everything downstream of the merge records (sm_100a).
//
// expand_rows      records -> packed microcanonical_statistics_dtype rows
//                  (percolate/hpc.py:57-70; the per-bond bookkeeping of
//                  hpc.py:249-307 re-expressed as a block-wide prefix scan)
#include <cstdlib>
#include "pz_common.cuh"
#include "pz_internal.h"

namespace pz {

// ---------------------------------------------------------------------------
// block-wide inclusive scan of the run statistics deltas
// ---------------------------------------------------------------------------
struct Delta {
    uint32_t c, mx;
    uint64_t s2, s3, s4;
};

__device__ __forceinline__ Delta delta_combine(const Delta &a, const Delta &b) {
    Delta r;
    r.c = a.c + b.c;
    r.mx = a.mx > b.mx ? a.mx : b.mx;
    r.s2 = a.s2 + b.s2; r.s3 = a.s3 + b.s3; r.s4 = a.s4 + b.s4;
    return r;
}
__device__ __forceinline__ Delta delta_shfl_up(const Delta &d, int k) {
    Delta r;
    r.c = __shfl_up_sync(0xffffffffu, d.c, k);
    r.mx = __shfl_up_sync(0xffffffffu, d.mx, k);
    r.s2 = __shfl_up_sync(0xffffffffu, d.s2, k);
    r.s3 = __shfl_up_sync(0xffffffffu, d.s3, k);
    r.s4 = __shfl_up_sync(0xffffffffu, d.s4, k);
    return r;
}
template <class RecT>
__device__ __forceinline__ Delta delta_of(RecT r) {
    Delta d{0, 0, 0, 0, 0};
    if (RecCodec<RecT>::valid(r)) {
        // w = a + b, t = a b:  w^2 - a^2 - b^2 = 2t,  w^3 - a^3 - b^3 = 3tw,
        // w^4 - a^4 - b^4 = 2t (2 w^2 - t)   (identities in Z, hence mod 2^64)
        const uint32_t a = RecCodec<RecT>::w_small(r), b = RecCodec<RecT>::w_large(r), w = a + b;
        const uint64_t t = (uint64_t)a * b, w2 = (uint64_t)w * w;
        d.c = 1; d.mx = w;
        d.s2 = 2 * t;
        d.s3 = 3 * t * w;
        d.s4 = 2 * t * (2 * w2 - t);
    }
    return d;
}

static constexpr int ROWS_THREADS = 256;

template <int BYTE_OFF>
__device__ __forceinline__ void put32(uint32_t (&w)[14], uint32_t v) {
    constexpr int i = BYTE_OFF / 4, s = (BYTE_OFF % 4) * 8;
    w[i] |= v << s;
    if constexpr (s != 0) w[i + 1] |= v >> ((32 - s) & 31);
}
template <int BYTE_OFF>
__device__ __forceinline__ void put64(uint32_t (&w)[14], uint64_t v) {
    put32<BYTE_OFF>(w, (uint32_t)v);
    put32<BYTE_OFF + 4>(w, (uint32_t)(v >> 32));
}

// one CTA per run; rows 0..M in chunks of ROWS_THREADS
template <class RecT, bool SPANNING>
__global__ void __launch_bounds__(ROWS_THREADS) expand_rows_kernel(StatsArgs a, uint8_t *rows)
{
    constexpr int RB = SPANNING ? 53 : 52;
    __shared__ Delta warp_tot[ROWS_THREADS / 32];
    __shared__ Delta carry;
    __shared__ __align__(16) uint8_t stage[ROWS_THREADS * RB + 8];

    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int run = blockIdx.x;
    const RecT *recs = reinterpret_cast<const RecT *>(a.recs) + (size_t)run * a.M;
    const int32_t *perm = a.perms ? a.perms + (size_t)run * a.M : nullptr;
    const uint32_t nspan = a.nspan[run];
    const int rows_total = a.M + 1;
    if (t == 0) carry = Delta{0u, 1u, (uint64_t)a.N, (uint64_t)a.N, (uint64_t)a.N};   // n = 0: c = 0, max = 1, S_k = N
    __syncthreads();

    for (int n0 = 0; n0 < rows_total; n0 += ROWS_THREADS) {
        const int n = n0 + t;
        const bool valid = n < rows_total;
        Delta d{0, 0, 0, 0, 0};
        if (valid && n >= 1) d = delta_of<RecT>(recs[n - 1]);
#pragma unroll
        for (int k = 1; k < 32; k <<= 1) {
            Delta o = delta_shfl_up(d, k);
            if (lane >= k) d = delta_combine(o, d);
        }
        if (lane == 31) warp_tot[warp] = d;
        __syncthreads();
        Delta pre = carry;
        for (int w = 0; w < warp; ++w) pre = delta_combine(pre, warp_tot[w]);
        d = delta_combine(pre, d);
        __syncthreads();
        if (t == ROWS_THREADS - 1) carry = d;

        const int nrows = min(ROWS_THREADS, rows_total - n0);
        const size_t goff = ((size_t)run * rows_total + n0) * RB;
        const int shift = (int)(goff & 3);
        if (valid) {
            uint32_t w[14];
#pragma unroll
            for (int i = 0; i < 14; ++i) w[i] = 0;
            const uint64_t x = d.mx, x2 = x * x;
            w[0] = (uint32_t)n;
            w[1] = (n >= 1 && perm) ? (uint32_t)perm[n - 1] : 0u;
            if (SPANNING) {
                w[2] = (uint32_t)n >= nspan ? 1u : 0u;
                put32<9>(w, d.mx);
                put64<13>(w, (uint64_t)a.N - 1 - d.c);
                put64<21>(w, (uint64_t)a.N - x);
                put64<29>(w, d.s2 - x2);
                put64<37>(w, d.s3 - x2 * x);
                put64<45>(w, d.s4 - x2 * x2);
            } else {
                put32<8>(w, d.mx);
                put64<12>(w, (uint64_t)a.N - 1 - d.c);
                put64<20>(w, (uint64_t)a.N - x);
                put64<28>(w, d.s2 - x2);
                put64<36>(w, d.s3 - x2 * x);
                put64<44>(w, d.s4 - x2 * x2);
            }
            uint8_t *dst = stage + shift + t * RB;
#pragma unroll
            for (int i = 0; i < RB; ++i) dst[i] = (uint8_t)(w[i / 4] >> ((i % 4) * 8));
        }
        __syncthreads();
        // coalesced copy-out: head bytes, aligned words, tail bytes
        {
            const int total = nrows * RB;
            const int head = min((4 - shift) & 3, total);
            uint8_t *g = rows + goff;
            const uint8_t *s = stage + shift;
            if (t < head) g[t] = s[t];
            const int nwords = (total - head) / 4;
            const uint32_t *s32 = reinterpret_cast<const uint32_t *>(s + head);
            uint32_t *g32 = reinterpret_cast<uint32_t *>(g + head);
            for (int i = t; i < nwords; i += ROWS_THREADS) g32[i] = s32[i];
            const int done = head + nwords * 4;
            if (t < total - done) g[done + t] = s[done + t];
        }
        __syncthreads();
    }
}


// ===========================================================================
// accumulate: per-n exact integer sums over runs (the inputs of
// _microcanonical_average_*, percolate/percolate.py:450-705).
//
// One lane owns one run and walks its records in n; a warp therefore advances
// 32 runs in lock step.  Records are staged through a padded shared-memory
// tile so that global reads are 128-byte coalesced and the per-lane walk is
// bank-conflict free.  For every row the 25 accumulator words of the 32 runs
// are summed with a recursive-halving exchange (31 64-bit shuffles; lane w
// ends up holding the warp total of word w) and added to the global
// accumulators with one 64-bit atomic per word per warp.
//
// Accumulator words per n (all uint64; 32-bit limbs so that word-wise integer
// sums over any number of partials -- warps here, GPUs in the all-reduce --
// are exact):
//   0        runs whose spanning cluster FIRST appears at n (delta form)
//   1        sum max           2,3    sum max^2      (lo32, hi32)
//   4        sum c             5,6    sum c^2        (c = merges so far)
//   7+6j..   j = 0,1,2 for moments[2+j] = m:
//            sum m (lo32, hi32), sum m^2 (four 32-bit limbs)
// A warp covers one segment of rows of its 32 runs, starting from the run
// states that checkpoint_kernel left at the segment boundary.
// ===========================================================================
static constexpr int ACC_WORDS = 25;

template <int HALF>
__device__ __forceinline__ void halve_step(uint64_t (&v)[32], int lane) {
    const bool up = lane & HALF;
#pragma unroll
    for (int i = 0; i < HALF; ++i) {
        const uint64_t send = up ? v[i] : v[i + HALF];
        const uint64_t keep = up ? v[i + HALF] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, HALF);
    }
}

// checkpoints: run state after every `every`-th row, by a block-wide scan
// (one CTA per run, CK_ITEMS consecutive records per thread and chunk) -- lets
// accumulate / canon_runs start anywhere in a run.  `every` must be a multiple
// of CK_ITEMS.
static constexpr int CK_ITEMS = 8;

template <class RecT>
__global__ void __launch_bounds__(ROWS_THREADS) checkpoint_kernel(StatsArgs a, RunState *ckpt,
                                                                   int every, int n_ckpt)
{
    __shared__ Delta warp_tot[ROWS_THREADS / 32];
    __shared__ Delta carry;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int run = blockIdx.x;
    const RecT *recs = reinterpret_cast<const RecT *>(a.recs) + (size_t)run * a.M;
    const int rows_total = a.M + 1;
    constexpr int CHUNK = ROWS_THREADS * CK_ITEMS;
    if (t == 0) carry = Delta{0u, 1u, (uint64_t)a.N, (uint64_t)a.N, (uint64_t)a.N};
    __syncthreads();
    for (int n0 = 0; n0 < rows_total; n0 += CHUNK) {
        const int first = n0 + t * CK_ITEMS;          // this thread's rows: first .. first+7
        RecT r[CK_ITEMS];
#pragma unroll
        for (int i = 0; i < CK_ITEMS; ++i) {
            const int n = first + i;
            r[i] = (n >= 1 && n < rows_total) ? __ldg(&recs[n - 1]) : (RecT)0;
        }
        const Delta d0 = delta_of<RecT>(r[0]);        // first row of the thread
        Delta d = d0;
#pragma unroll
        for (int i = 1; i < CK_ITEMS; ++i) d = delta_combine(d, delta_of<RecT>(r[i]));
        const Delta mine = d;
#pragma unroll
        for (int k = 1; k < 32; k <<= 1) {
            Delta o = delta_shfl_up(d, k);
            if (lane >= k) d = delta_combine(o, d);
        }
        if (lane == 31) warp_tot[warp] = d;
        __syncthreads();
        Delta pre = carry;
        for (int w = 0; w < warp; ++w) pre = delta_combine(pre, warp_tot[w]);
        // exclusive prefix of this thread = pre + (inclusive - mine): recombine from lane-1
        Delta excl = delta_shfl_up(d, 1);
        if (lane == 0) excl = Delta{0, 0, 0, 0, 0};
        pre = delta_combine(pre, excl);
        const Delta incl = delta_combine(pre, mine);
        __syncthreads();
        if (t == ROWS_THREADS - 1) carry = incl;
        if (first < rows_total && (first % every) == 0) {
            const Delta s0 = delta_combine(pre, d0);  // state after row `first`
            RunState st;
            st.c = s0.c; st.mx = s0.mx; st.s2 = s0.s2; st.s3 = s0.s3; st.s4 = s0.s4;
            ckpt[(size_t)run * n_ckpt + first / every] = st;
        }
        __syncthreads();
    }
}

// The same checkpoints with ONE WARP PER RUN: a lane takes 8 consecutive records (two 16-byte loads
// when the rows of the record array are aligned), the warp scans 256 records per step and carries
// the running state in registers -- no shared memory, no CTA barriers, no cross-warp prefix.  A
// batch has thousands of runs, so a warp per run still fills the GPU.  The state after row k (the
// records before k applied) is the exclusive prefix of the lane that starts at record k.
__device__ __forceinline__ Delta delta_shfl(const Delta &d, int src) {
    Delta r;
    r.c = __shfl_sync(0xffffffffu, d.c, src);
    r.mx = __shfl_sync(0xffffffffu, d.mx, src);
    r.s2 = __shfl_sync(0xffffffffu, d.s2, src);
    r.s3 = __shfl_sync(0xffffffffu, d.s3, src);
    r.s4 = __shfl_sync(0xffffffffu, d.s4, src);
    return r;
}

template <class RecT>
__device__ __forceinline__ void load8_records(const RecT *recs, int M, int k0, bool vec, RecT (&r)[CK_ITEMS])
{
    if constexpr (sizeof(RecT) == 4) {
        if (vec && k0 + CK_ITEMS <= M) {
            const uint4 *p = reinterpret_cast<const uint4 *>(recs + k0);
            const uint4 x = __ldg(p), y = __ldg(p + 1);
            r[0] = x.x; r[1] = x.y; r[2] = x.z; r[3] = x.w;
            r[4] = y.x; r[5] = y.y; r[6] = y.z; r[7] = y.w;
            return;
        }
    }
#pragma unroll
    for (int i = 0; i < CK_ITEMS; ++i) r[i] = (k0 + i < M) ? __ldg(&recs[k0 + i]) : (RecT)0;
}

template <class RecT>
__global__ void __launch_bounds__(128) checkpoint_warp_kernel(StatsArgs a, RunState *ckpt, int every,
                                                              int n_ckpt)
{
    static_assert(CK_ITEMS == 8, "a lane takes 8 records");
    const int lane = threadIdx.x & 31;
    const int run = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (run >= a.R) return;                                   // the whole warp leaves
    const int M = a.M;
    const RecT *recs = reinterpret_cast<const RecT *>(a.recs) + (size_t)run * M;
    RunState *out = ckpt + (size_t)run * n_ckpt;
    const bool vec = sizeof(RecT) == 4 && (M & 3) == 0;      // every run's records start 16-byte aligned
    constexpr int STEP = 32 * CK_ITEMS;
    Delta carry{0u, 1u, (uint64_t)a.N, (uint64_t)a.N, (uint64_t)a.N};      // the state at row 0
    RecT nxt[CK_ITEMS];
    load8_records<RecT>(recs, M, CK_ITEMS * lane, vec, nxt);
    for (int base = 0; base <= M; base += STEP) {
        const int k0 = base + CK_ITEMS * lane;                // this lane: records k0 .. k0+7
        RecT r[CK_ITEMS];
#pragma unroll
        for (int i = 0; i < CK_ITEMS; ++i) r[i] = nxt[i];
        if (base + STEP <= M) load8_records<RecT>(recs, M, k0 + STEP, vec, nxt);   // the next step's records
        Delta d = delta_of<RecT>(r[0]);
#pragma unroll
        for (int i = 1; i < CK_ITEMS; ++i) d = delta_combine(d, delta_of<RecT>(r[i]));
#pragma unroll
        for (int k = 1; k < 32; k <<= 1) {
            const Delta o = delta_shfl_up(d, k);
            if (lane >= k) d = delta_combine(o, d);
        }
        Delta excl = delta_shfl_up(d, 1);
        if (lane == 0) excl = Delta{0, 0, 0, 0, 0};
        if (k0 <= M && (k0 % every) == 0) {
            const Delta pre = delta_combine(carry, excl);     // state after row k0
            RunState st;
            st.c = pre.c; st.mx = pre.mx; st.s2 = pre.s2; st.s3 = pre.s3; st.s4 = pre.s4;
            out[k0 / every] = st;
        }
        carry = delta_combine(carry, delta_shfl(d, 31));
    }
}

// PZ_CKPT_WARP=0: the block-scan kernel (one CTA per run)
static int ckpt_warp_mode()
{
    static int mode = -1;
    if (mode < 0) {
        const char *e = getenv("PZ_CKPT_WARP");
        mode = e ? (atoi(e) != 0) : 1;
    }
    return mode;
}

cudaError_t launch_checkpoints(const StatsArgs &a, RunState *ckpt, int every, int n_ckpt,
                               cudaStream_t s)
{
    if (a.R <= 0) return cudaSuccess;
    if (ckpt_warp_mode() && every % CK_ITEMS == 0) {
        const int grid = (a.R + 3) / 4;
        if (a.rec64) checkpoint_warp_kernel<uint64_t><<<grid, 128, 0, s>>>(a, ckpt, every, n_ckpt);
        else checkpoint_warp_kernel<uint32_t><<<grid, 128, 0, s>>>(a, ckpt, every, n_ckpt);
        return cudaGetLastError();
    }
    if (a.rec64) checkpoint_kernel<uint64_t><<<a.R, ROWS_THREADS, 0, s>>>(a, ckpt, every, n_ckpt);
    else checkpoint_kernel<uint32_t><<<a.R, ROWS_THREADS, 0, s>>>(a, ckpt, every, n_ckpt);
    return cudaGetLastError();
}

// one warp = (32 runs) x (one segment of `seg` rows starting at a checkpoint)
template <class RecT>
__global__ void __launch_bounds__(256) accumulate_kernel(StatsArgs a, unsigned long long *acc,
                                                          const RunState *ckpt, int seg, int n_ckpt)
{
    constexpr int TILE = sizeof(RecT) == 4 ? 32 : 16;
    __shared__ RecT tile_all[8][32][TILE + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    RecT (*tile)[TILE + 1] = tile_all[warp];
    const long long wid = (long long)blockIdx.x * 8 + warp;
    const int sgi = (int)(wid % n_ckpt);                 // segment index
    const int run0 = (int)(wid / n_ckpt) * 32;
    if (run0 >= a.R) return;
    const int run = run0 + lane;
    const bool live = run < a.R;
    const RecT *recs = reinterpret_cast<const RecT *>(a.recs);
    const int M = a.M;
    const int row_lo = sgi * seg, row_hi = min(M, row_lo + seg - 1);   // rows of this segment

    RunState st;
    st.init((uint32_t)a.N);
    if (live) st = ckpt[(size_t)run * n_ckpt + sgi];     // state AFTER row_lo
    if (sgi == 0 && live && a.spanning) {
        const uint32_t ns = a.nspan[run];
        if (ns != NSPAN_NEVER) atomicAdd(&acc[(size_t)ns * ACC_WORDS + 0], 1ull);
    }

    auto emit = [&](int n) {
        uint64_t v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0;
        if (live) {
            const uint64_t x = st.mx, x2 = x * x, c = st.c, c2 = c * c;
            v[1] = x;
            v[2] = x2 & 0xffffffffu; v[3] = x2 >> 32;
            v[4] = c;
            v[5] = c2 & 0xffffffffu; v[6] = c2 >> 32;
            const uint64_t m[3] = {st.s2 - x2, st.s3 - x2 * x, st.s4 - x2 * x2};
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const uint64_t lo = m[k] * m[k], hi = __umul64hi(m[k], m[k]);
                v[7 + 6 * k + 0] = m[k] & 0xffffffffu;
                v[7 + 6 * k + 1] = m[k] >> 32;
                v[7 + 6 * k + 2] = lo & 0xffffffffu;
                v[7 + 6 * k + 3] = lo >> 32;
                v[7 + 6 * k + 4] = hi & 0xffffffffu;
                v[7 + 6 * k + 5] = hi >> 32;
            }
        }
        halve_step<16>(v, lane);
        halve_step<8>(v, lane);
        halve_step<4>(v, lane);
        halve_step<2>(v, lane);
        halve_step<1>(v, lane);
        if (lane >= 1 && lane < ACC_WORDS && v[0])
            atomicAdd(&acc[(size_t)n * ACC_WORDS + lane], (unsigned long long)v[0]);
    };

    emit(row_lo);
    for (int row = row_lo; row < row_hi; row += TILE) {
        // records row .. row+TILE-1 produce rows row+1 .. row+TILE
        __syncwarp();
        if (lane < TILE) {
            const int idx = row + lane;
            for (int rr = 0; rr < 32; ++rr) {
                RecT v = 0;
                if (run0 + rr < a.R && idx < M) v = __ldcs(&recs[(size_t)(run0 + rr) * M + idx]);
                tile[rr][lane] = v;
            }
        }
        __syncwarp();
        const int cnt = min(TILE, row_hi - row);
        for (int j = 0; j < cnt; ++j) {
            const RecT r = tile[lane][j];
            if (RecCodec<RecT>::valid(r)) st.merge(RecCodec<RecT>::w_small(r), RecCodec<RecT>::w_large(r));
            emit(row + 1 + j);
        }
    }
}

cudaError_t launch_accumulate(const StatsArgs &a, unsigned long long *acc, const RunState *ckpt,
                              int seg, int n_ckpt, cudaStream_t s)
{
    if (a.R <= 0) return cudaSuccess;
    const long long warps = (long long)((a.R + 31) / 32) * n_ckpt;
    const int grid = (int)((warps + 7) / 8);
    if (a.rec64) accumulate_kernel<uint64_t><<<grid, 256, 0, s>>>(a, acc, ckpt, seg, n_ckpt);
    else accumulate_kernel<uint32_t><<<grid, 256, 0, s>>>(a, acc, ckpt, seg, n_ckpt);
    return cudaGetLastError();
}

// ===========================================================================
// accumulate, tile form (the one the fused path uses): one LANE owns two
// consecutive rows of a 64-row tile, one warp walks the runs of a group for that
// tile.  Per run: two coalesced record loads per lane, one broadcast load of the
// run state at the tile boundary (checkpoint every 64 rows), a warp scan of the
// 2-row deltas, then the row statistics are added to REGISTER accumulators
// (carry chains, 64/128/192 bits).  Only after the last run of the group are
// the totals split into the 32-bit-limb words of the accumulator block and
// added with one 64-bit atomic per (row, word) -- no shared memory, no per-run
// exchange between lanes beyond the scan.
// ===========================================================================
static constexpr int ACC_TILE = 64;            // rows per warp tile = checkpoint spacing
static constexpr int ACC_GROUP = 512;          // runs per warp

struct U128 { uint64_t lo, hi; };
struct U192 { uint64_t w0, w1, w2; };
__device__ __forceinline__ void add128(U128 &a, uint64_t b) {
    asm("add.cc.u64 %0, %0, %2;\n\taddc.u64 %1, %1, 0;" : "+l"(a.lo), "+l"(a.hi) : "l"(b));
}
__device__ __forceinline__ void add192(U192 &a, uint64_t blo, uint64_t bhi) {
    asm("add.cc.u64 %0, %0, %3;\n\taddc.cc.u64 %1, %1, %4;\n\taddc.u64 %2, %2, 0;"
        : "+l"(a.w0), "+l"(a.w1), "+l"(a.w2) : "l"(blo), "l"(bhi));
}

struct RowAcc {
    uint64_t mx, c;            // sum max, sum c          (< 2^29 * runs of a group)
    U128 mx2, c2;              // sum max^2, sum c^2
    U128 m[3];                 // sum moments[2..4]
    U192 mm[3];                // sum moments[2..4]^2
    __device__ __forceinline__ void clear() {
        mx = c = 0; mx2 = c2 = U128{0, 0};
#pragma unroll
        for (int k = 0; k < 3; ++k) { m[k] = U128{0, 0}; mm[k] = U192{0, 0, 0}; }
    }
    __device__ __forceinline__ void add(uint32_t sc, uint32_t smx, uint64_t s2, uint64_t s3, uint64_t s4) {
        const uint64_t x = smx, x2 = x * x, cc = sc;
        mx += x; c += cc;
        add128(mx2, x2);
        add128(c2, cc * cc);
        const uint64_t v[3] = {s2 - x2, s3 - x2 * x, s4 - x2 * x2};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            add128(m[k], v[k]);
            add192(mm[k], v[k] * v[k], __umul64hi(v[k], v[k]));
        }
    }
    // split into the limb words of the accumulator block and add them
    __device__ __forceinline__ void flush(unsigned long long *w) const {
        auto put = [&](int i, uint64_t v) { if (v) atomicAdd(&w[i], (unsigned long long)v); };
        put(1, mx);
        put(2, mx2.lo & 0xffffffffu); put(3, (mx2.lo >> 32) | (mx2.hi << 32));
        put(4, c);
        put(5, c2.lo & 0xffffffffu); put(6, (c2.lo >> 32) | (c2.hi << 32));
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            put(7 + 6 * k + 0, m[k].lo & 0xffffffffu);
            put(7 + 6 * k + 1, (m[k].lo >> 32) | (m[k].hi << 32));
            put(7 + 6 * k + 2, mm[k].w0 & 0xffffffffu);
            put(7 + 6 * k + 3, mm[k].w0 >> 32);
            put(7 + 6 * k + 4, mm[k].w1 & 0xffffffffu);
            put(7 + 6 * k + 5, (mm[k].w1 >> 32) | (mm[k].w2 << 32));
        }
    }
};

template <class RecT>
__global__ void __launch_bounds__(128, 3) accumulate_tiles_kernel(StatsArgs a, unsigned long long *acc,
                                                                  const RunState *ckpt, int n_ckpt,
                                                                  int n_tiles)
{
    const int lane = threadIdx.x & 31;
    const long long wid = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
    const int tile = (int)(wid % n_tiles);
    const int run_lo = (int)(wid / n_tiles) * ACC_GROUP;
    if (run_lo >= a.R) return;
    const int run_hi = min(a.R, run_lo + ACC_GROUP);
    const int M = a.M;
    const RecT *recs = reinterpret_cast<const RecT *>(a.recs);
    // this lane: records i0, i0 + 1 -> rows i0 + 1, i0 + 2
    const int i0 = tile * ACC_TILE + 2 * lane;
    const bool v0 = i0 < M, v1 = i0 + 1 < M;

    if (tile == 0) {
        if (a.spanning)
            for (int r = run_lo + lane; r < run_hi; r += 32) {
                const uint32_t ns = a.nspan[r];
                if (ns != NSPAN_NEVER) atomicAdd(&acc[(size_t)ns * ACC_WORDS + 0], 1ull);
            }
        if (lane == 0) {
            // row 0 is the same for every run (max = 1, c = 0, moments = N - 1):
            // limb * count is exact in the limb-word format
            const unsigned long long cnt = (unsigned long long)(run_hi - run_lo);
            const uint64_t v = (uint64_t)a.N - 1, vv = v * v;
            atomicAdd(&acc[1], cnt);
            atomicAdd(&acc[2], cnt);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                atomicAdd(&acc[7 + 6 * k + 0], (unsigned long long)(v & 0xffffffffu) * cnt);
                atomicAdd(&acc[7 + 6 * k + 1], (unsigned long long)(v >> 32) * cnt);
                atomicAdd(&acc[7 + 6 * k + 2], (unsigned long long)(vv & 0xffffffffu) * cnt);
                atomicAdd(&acc[7 + 6 * k + 3], (unsigned long long)(vv >> 32) * cnt);
            }
        }
    }

    RowAcc A, B;
    A.clear(); B.clear();

    auto load = [&](int run, RecT &r0, RecT &r1, RunState &st) {
        const RecT *p = recs + (size_t)run * M + i0;
        r0 = v0 ? __ldcs(p) : (RecT)0;
        r1 = v1 ? __ldcs(p + 1) : (RecT)0;
        const uint4 *q = reinterpret_cast<const uint4 *>(ckpt + (size_t)run * n_ckpt + tile);
        const uint4 lo = __ldg(q), hi = __ldg(q + 1);
        st.c = lo.x; st.mx = lo.y;
        st.s2 = ((uint64_t)lo.w << 32) | lo.z;
        st.s3 = ((uint64_t)hi.y << 32) | hi.x;
        st.s4 = ((uint64_t)hi.w << 32) | hi.z;
    };

    RecT n0 = 0, n1 = 0;
    RunState nst;
    nst.init((uint32_t)a.N);
    load(run_lo, n0, n1, nst);
    for (int run = run_lo; run < run_hi; ++run) {
        const RecT r0 = n0, r1 = n1;
        const RunState base = nst;
        if (run + 1 < run_hi) load(run + 1, n0, n1, nst);      // prefetch the next run

        const Delta d0 = delta_of<RecT>(r0), d1 = delta_of<RecT>(r1);
        Delta d = delta_combine(d0, d1);
#pragma unroll
        for (int k = 1; k < 32; k <<= 1) {
            const Delta o = delta_shfl_up(d, k);
            if (lane >= k) d = delta_combine(o, d);
        }
        Delta ex = delta_shfl_up(d, 1);                        // exclusive prefix of this lane
        if (lane == 0) ex = Delta{0, 0, 0, 0, 0};
        // state after row i0 + 1
        const uint32_t ca = base.c + ex.c + d0.c;
        const uint32_t xa = max(max(base.mx, ex.mx), d0.mx);
        const uint64_t s2a = base.s2 + ex.s2 + d0.s2, s3a = base.s3 + ex.s3 + d0.s3,
                       s4a = base.s4 + ex.s4 + d0.s4;
        if (v0) A.add(ca, xa, s2a, s3a, s4a);
        if (v1) B.add(ca + d1.c, max(xa, d1.mx), s2a + d1.s2, s3a + d1.s3, s4a + d1.s4);
    }
    if (v0) A.flush(acc + (size_t)(i0 + 1) * ACC_WORDS);
    if (v1) B.flush(acc + (size_t)(i0 + 2) * ACC_WORDS);
}

cudaError_t launch_accumulate_tiles(const StatsArgs &a, unsigned long long *acc, const RunState *ckpt,
                                    int every, int n_ckpt, cudaStream_t s)
{
    if (a.R <= 0) return cudaSuccess;
    if (every != ACC_TILE) return cudaErrorInvalidValue;
    const int n_tiles = a.M > 0 ? (a.M + ACC_TILE - 1) / ACC_TILE : 1;
    const long long warps = (long long)((a.R + ACC_GROUP - 1) / ACC_GROUP) * n_tiles;
    const int grid = (int)((warps + 3) / 4);
    if (a.rec64) accumulate_tiles_kernel<uint64_t><<<grid, 128, 0, s>>>(a, acc, ckpt, n_ckpt, n_tiles);
    else accumulate_tiles_kernel<uint32_t><<<grid, 128, 0, s>>>(a, acc, ckpt, n_ckpt, n_tiles);
    return cudaGetLastError();
}

// ===========================================================================
// micro_finalize: exact sums -> float64 mean and unbiased variance per n.
// var = (R * sum x^2 - (sum x)^2) / (R (R-1)) is formed in 256-bit integer
// arithmetic, so it is exactly 0 when all runs agree (the reference's
// ``if std:`` branches, percolate/percolate.py:621-633, 691-703) and carries
// no cancellation error otherwise.
// ===========================================================================
struct U256 {
    uint64_t w[4];
};
__device__ __forceinline__ U256 u256_zero() { U256 r; r.w[0] = r.w[1] = r.w[2] = r.w[3] = 0; return r; }
__device__ __forceinline__ void u256_add_shifted(U256 &x, uint64_t v, int bit) {
    // x += v << bit   (bit multiple of 32)
    const int limb = bit >> 6, sh = bit & 63;
    uint64_t lo = v << sh, hi = sh ? (v >> (64 - sh)) : 0;
    unsigned carry = 0;
    for (int i = limb; i < 4; ++i) {
        const uint64_t add = (i == limb ? lo : (i == limb + 1 ? hi : 0));
        const uint64_t s1 = x.w[i] + add;
        const unsigned c1 = s1 < add;
        const uint64_t s2 = s1 + carry;
        const unsigned c2 = s2 < s1;
        x.w[i] = s2;
        carry = c1 | c2;
    }
}
__device__ __forceinline__ U256 u256_mul64(const U256 &x, uint64_t m) {
    U256 r = u256_zero();
    uint64_t carry = 0;
    for (int i = 0; i < 4; ++i) {
        const uint64_t lo = x.w[i] * m, hi = __umul64hi(x.w[i], m);
        const uint64_t s = lo + carry;
        r.w[i] = s;
        carry = hi + (s < lo);
    }
    return r;
}
__device__ __forceinline__ U256 u256_sqr128(uint64_t a0, uint64_t a1) {
    // (a0 + a1 2^64)^2
    U256 r = u256_zero();
    r.w[0] = a0 * a0; r.w[1] = __umul64hi(a0, a0);
    const uint64_t clo = a0 * a1, chi = __umul64hi(a0, a1);
    // 2 * cross << 64
    U256 cross = u256_zero();
    cross.w[1] = clo << 1; cross.w[2] = (chi << 1) | (clo >> 63); cross.w[3] = chi >> 63;
    U256 top = u256_zero();
    top.w[2] = a1 * a1; top.w[3] = __umul64hi(a1, a1);
    unsigned carry = 0;
    for (int i = 0; i < 4; ++i) {
        uint64_t s = r.w[i] + cross.w[i]; unsigned c1 = s < cross.w[i];
        uint64_t s2 = s + top.w[i]; unsigned c2 = s2 < top.w[i];
        uint64_t s3 = s2 + carry; unsigned c3 = s3 < s2;
        r.w[i] = s3; carry = c1 + c2 + c3;
    }
    return r;
}
__device__ __forceinline__ U256 u256_sub(const U256 &a, const U256 &b) {
    U256 r; unsigned borrow = 0;
    for (int i = 0; i < 4; ++i) {
        const uint64_t d = a.w[i] - b.w[i]; const unsigned b1 = a.w[i] < b.w[i];
        const uint64_t d2 = d - borrow; const unsigned b2 = d < borrow;
        r.w[i] = d2; borrow = b1 | b2;
    }
    return r;
}
__device__ __forceinline__ double u256_to_double(const U256 &x) {
    int top = 3;
    while (top > 0 && x.w[top] == 0) --top;
    if (top == 0) return (double)x.w[0];
    const double hi = (double)x.w[top], lo = (double)x.w[top - 1];
    return ldexp(hi * 18446744073709551616.0 + lo, 64 * (top - 1));
}
// sum x as (lo32-word, hi32-word) -> 128-bit
__device__ __forceinline__ void sum128(uint64_t wlo, uint64_t whi, uint64_t &a0, uint64_t &a1) {
    U256 t = u256_zero();
    u256_add_shifted(t, wlo, 0);
    u256_add_shifted(t, whi, 32);
    a0 = t.w[0]; a1 = t.w[1];
}
__device__ __forceinline__ double var_exact(uint64_t a0, uint64_t a1, const U256 &B, uint64_t R) {
    if (R < 2) return nan("");
    const U256 RB = u256_mul64(B, R);
    const U256 A2 = u256_sqr128(a0, a1);
    const U256 D = u256_sub(RB, A2);
    return u256_to_double(D) / ((double)R * (double)(R - 1));
}
__device__ __forceinline__ double mean_exact(uint64_t a0, uint64_t a1, uint64_t R) {
    U256 t = u256_zero(); t.w[0] = a0; t.w[1] = a1;
    return u256_to_double(t) / (double)R;
}

// mean[7][M+1]: spanning count, max, moments 0..4 ; var[6][M+1]: max, moments 0..4
__global__ void micro_finalize_kernel(int32_t N, int32_t M, unsigned long long R,
                                      const unsigned long long *acc,
                                      const unsigned long long *span_cum, double *mean, double *var)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n > M) return;
    const unsigned long long *w = acc + (size_t)n * ACC_WORDS;
    const size_t S = (size_t)M + 1;
    mean[0 * S + n] = (double)span_cum[n];
    // max
    {
        uint64_t b0, b1; sum128(w[2], w[3], b0, b1);
        U256 B = u256_zero(); B.w[0] = b0; B.w[1] = b1;
        mean[1 * S + n] = mean_exact(w[1], 0, R);
        const double v = var_exact(w[1], 0, B, R);
        var[0 * S + n] = v;
        // moments[1] = N - max  (percolate/hpc.py:214, 283-305)
        const uint64_t s1 = (uint64_t)N * R - w[1];
        mean[3 * S + n] = mean_exact(s1, 0, R);
        var[2 * S + n] = v;
    }
    // moments[0] = N - 1 - c
    {
        uint64_t b0, b1; sum128(w[5], w[6], b0, b1);
        U256 B = u256_zero(); B.w[0] = b0; B.w[1] = b1;
        const uint64_t s0 = (uint64_t)(N - 1) * R - w[4];
        mean[2 * S + n] = mean_exact(s0, 0, R);
        var[1 * S + n] = var_exact(w[4], 0, B, R);
    }
    for (int k = 0; k < 3; ++k) {
        const unsigned long long *q = w + 7 + 6 * k;
        uint64_t a0, a1; sum128(q[0], q[1], a0, a1);
        U256 B = u256_zero();
        u256_add_shifted(B, q[2], 0);
        u256_add_shifted(B, q[3], 32);
        u256_add_shifted(B, q[4], 64);
        u256_add_shifted(B, q[5], 96);
        mean[(4 + k) * S + n] = mean_exact(a0, a1, R);
        var[(3 + k) * S + n] = var_exact(a0, a1, B, R);
    }
}

// inclusive prefix sum of word 0 (delta form -> runs spanning at n); one CTA
__global__ void span_cumsum_kernel(int32_t M, const unsigned long long *acc, unsigned long long *out)
{
    __shared__ unsigned long long warp_tot[32];
    __shared__ unsigned long long carry;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    if (t == 0) carry = 0;
    __syncthreads();
    for (int n0 = 0; n0 <= M; n0 += 1024) {
        const int n = n0 + t;
        unsigned long long v = n <= M ? acc[(size_t)n * ACC_WORDS] : 0ull;
        for (int k = 1; k < 32; k <<= 1) {
            const unsigned long long o = __shfl_up_sync(0xffffffffu, v, k);
            if (lane >= k) v += o;
        }
        if (lane == 31) warp_tot[warp] = v;
        __syncthreads();
        unsigned long long pre = carry;
        for (int w = 0; w < warp; ++w) pre += warp_tot[w];
        v += pre;
        if (n <= M) out[n] = v;
        __syncthreads();
        if (t == 1023) carry = v;
        __syncthreads();
    }
}

cudaError_t launch_micro_finalize(int32_t N, int32_t M, int64_t runs, const unsigned long long *acc,
                                  unsigned long long *span_cum, double *mean, double *var,
                                  cudaStream_t s)
{
    span_cumsum_kernel<<<1, 1024, 0, s>>>(M, acc, span_cum);
    const int grid = (M + 1 + 127) / 128;
    micro_finalize_kernel<<<grid, 128, 0, s>>>(N, M, (unsigned long long)runs, acc, span_cum, mean, var);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// micro_arrays: the per-n arrays of microcanonical_averages_arrays
// (percolate/percolate.py:968-1064) from mean[7][S] / var[6][S]: sample mean and
// Student-t interval  t * std / sqrt(runs) + mean  (percolate/percolate.py:613-635,
// 681-705; (mean, mean) when the sample std is zero), each divided by `norm`
// (the number of sites, percolate/percolate.py:1056-1060).  Every operation is a
// separately rounded IEEE double operation in the order numpy performs them on the
// host, so the arrays are bit-identical to the host evaluation they replace.
// out: k[S] | max[S] | max_ci[S][2] | moments_ci[5][S][2] | moments[5][S]   (the interval
// blocks start at even offsets, so that their double2 stores are aligned for odd S too)
// ---------------------------------------------------------------------------
__global__ void micro_arrays_kernel(int32_t M, double sqrt_runs, double t_lo, double t_hi, double norm,
                                    const double *mean, const double *var, double *out)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n > M) return;
    const size_t S = (size_t)M + 1;
    out[n] = mean[n];                                       // runs spanning at n
    double *mx = out + S, *mom = out + 14 * S;
    double2 *mx_ci = reinterpret_cast<double2 *>(out + 2 * S);
    double2 *mom_ci = reinterpret_cast<double2 *>(out + 4 * S);
#pragma unroll
    for (int q = 0; q < 6; ++q) {                           // 0: max, 1..5: moments[0..4]
        const double m = mean[(size_t)(1 + q) * S + n];
        const double sd = sqrt(var[(size_t)q * S + n]);
        const double scale = __ddiv_rn(sd, sqrt_runs);
        double lo = __dadd_rn(__dmul_rn(t_lo, scale), m), hi = __dadd_rn(__dmul_rn(t_hi, scale), m);
        if (sd == 0.0) { lo = m; hi = m; }
        const double2 ci = make_double2(__ddiv_rn(lo, norm), __ddiv_rn(hi, norm));
        if (q == 0) { mx[n] = __ddiv_rn(m, norm); mx_ci[n] = ci; }
        else { mom[(size_t)(q - 1) * S + n] = __ddiv_rn(m, norm); mom_ci[(size_t)(q - 1) * S + n] = ci; }
    }
}

cudaError_t launch_micro_arrays(int32_t M, int64_t runs, double t_lo, double t_hi, double norm,
                                const double *mean, const double *var, double *out, cudaStream_t s)
{
    const int grid = (M + 1 + 127) / 128;
    micro_arrays_kernel<<<grid, 128, 0, s>>>(M, sqrt((double)runs), t_lo, t_hi, norm, mean, var, out);
    return cudaGetLastError();
}

cudaError_t launch_expand_rows(const StatsArgs &a, uint8_t *rows, cudaStream_t s)
{
    if (a.R <= 0) return cudaSuccess;
    if (a.rec64) {
        if (a.spanning) expand_rows_kernel<uint64_t, true><<<a.R, ROWS_THREADS, 0, s>>>(a, rows);
        else expand_rows_kernel<uint64_t, false><<<a.R, ROWS_THREADS, 0, s>>>(a, rows);
    } else {
        if (a.spanning) expand_rows_kernel<uint32_t, true><<<a.R, ROWS_THREADS, 0, s>>>(a, rows);
        else expand_rows_kernel<uint32_t, false><<<a.R, ROWS_THREADS, 0, s>>>(a, rows);
    }
    return cudaGetLastError();
}

}  // namespace pz
