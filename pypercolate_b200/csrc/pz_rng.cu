// pz_rng.cu -- bond orders generated on the device (sm_100a).
//
// The reference draws the bond order of a run on the host,
// ``RandomState(seed).permutation(M)`` (percolate/hpc.py:195,206).  Two device
// replacements (and two more of this package's own, further down):
//
//  * perm_mt19937: NumPy's legacy stream reproduced bit for bit -- MT19937
//    seeded by init_genrand(seed), then the legacy shuffle (for i = M-1..1:
//    j = masked-rejection draw in [0, i]; swap).  One WARP per run
//    (perm_warp_kernel): the twister state lives in registers across the warp,
//    32 shuffle steps are drawn and applied together.
//
//  * perm_philox: counter-based Philox4x32-10.  A uniform permutation is built
//    in two exact steps (Rao-Sandelius): every bond draws one of B buckets
//    uniformly and is placed by a counting sort whose buckets are kept in
//    ascending bond order (so the result does not depend on thread timing),
//    then every bucket is shuffled by Fisher-Yates in
//    shared memory with unbiased (Lemire) bounded draws.  One CTA per run.
//    Counters: bucket draws (i >> 2, 0, 0, 0) word i & 3; Fisher-Yates draw of
//    step k of bucket b: word k & 3 of (k >> 2, b, 0, 1), on the (rare) Lemire
//    rejection words 0.. of (k, b, attempt >= 1, 2).  Key: (seed, 0x50455243).
//    Every Philox call therefore serves four bonds.  (B grows with M up to 2^17: the mean bucket
//    stays near 64 bonds; M <= 2^24.)
//    oracle/pz_oracle.c restates this algorithm on the CPU for bit-exact tests.
//
//  * perm_feistel: the bond order as a keyed BIJECTION of [0, M): position n ->
//    bond pi_seed(n), evaluated independently per position -- no scatter, no
//    atomics, no shared memory, one coalesced write.  pi is an alternating
//    unbalanced Feistel network over the 2^k >= M points (k = ceil(log2 M),
//    halves of k/2 and k - k/2 bits that swap roles every round), 20 rounds:
//    with R = the low k - k/2 bits and L = the rest, f = hi ^ lo of the 64-bit
//    product (R ^ key_i) * 0xD2511F53 (the Philox multiplier),
//    x' = R << (k/2) | (L ^ f[k - k/2 .. k)); round keys = Philox4x32-10 words of
//    counter (i >> 2, 0, 0, 3) under key (seed, 'PERC'); points that land at or
//    above M are walked on (cycle walking) until they fall below M, which keeps
//    the map a bijection of [0, M).  A pseudo-random permutation family rather
//    than an exact uniform shuffle; tests/test_oracle.py checks it against the
//    uniform distribution over ALL permutations on small domains and by order
//    statistics on large ones, tests/test_gpu_parity.py against the reference's
//    confidence intervals.
#include <cstdlib>
#include <algorithm>
#include "pz_common.cuh"
#include "pz_internal.h"

namespace pz {

// ---------------------------------------------------------------------------
// Philox4x32-10
// ---------------------------------------------------------------------------
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t k0, uint32_t k1, uint32_t c0,
                                                       uint32_t c1, uint32_t c2, uint32_t c3,
                                                       uint32_t out[4])
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static constexpr uint32_t PHILOX_KEY1 = 0x50455243u;   // 'PERC'
#ifndef PZ_PH_THREADS
#define PZ_PH_THREADS 256
#endif
static constexpr int PH_THREADS = PZ_PH_THREADS;
static constexpr int PH_WARPS = PH_THREADS / 32;

// the four bucket words of bonds 4g .. 4g+3
__device__ __forceinline__ void philox_buckets4(uint32_t seed, uint32_t g, uint32_t (&o)[4])
{
    philox4x32_10(seed, PHILOX_KEY1, g, 0u, 0u, 0u, o);
}

// unbiased j in [0, k] (Lemire) from word k & 3 of counter (k >> 2, bucket, 0, 1), which the caller
// has selected; on the (rare) rejection words 0.. of (k, bucket, attempt >= 1, 2)
__device__ __forceinline__ uint32_t philox_bounded_w(uint32_t seed, uint32_t bucket, uint32_t k, uint32_t w)
{
    const uint32_t range = k + 1u;
    uint64_t m = (uint64_t)w * range;
    if ((uint32_t)m >= range) return (uint32_t)(m >> 32);
    const uint32_t thresh = (0u - range) % range;
    if ((uint32_t)m >= thresh) return (uint32_t)(m >> 32);
    for (uint32_t attempt = 1;; ++attempt) {
        uint32_t r[4];
        philox4x32_10(seed, PHILOX_KEY1, k, bucket, attempt, 2u, r);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            m = (uint64_t)r[q] * range;
            if ((uint32_t)m >= thresh) return (uint32_t)(m >> 32);
        }
    }
}

// ---------------------------------------------------------------------------
// perm_philox: Rao-Sandelius in two levels.  Every bond draws one of B = 2^lb buckets (mean
// bucket about 64 bonds, lb <= 17); the high bits of the bucket number name a SUPER-bucket of 128
// buckets (about 8000 bonds), the low 7 bits a bucket inside it.
//   A  per-warp counts of the super-buckets (one Philox call per four bonds), their starts;
//   B  every warp appends its bonds, in order, to its own segment of each super-bucket:
//      sequential streams, whose sectors fill within a few iterations and combine in the L2 (a
//      direct scatter to thousands of buckets leaves half-written sectors that do not survive
//      there with hundreds of runs in flight: that version was bound by read-modify-write
//      traffic to DRAM).  An entry carries its bucket: bond | low bits << 24;
//   C  per super-bucket: per-warp counts of its 128 buckets and their starts, placement into
//      shared memory in almost ascending order (warp after warp; only the entries of one warp
//      instruction may swap), one thread per bucket restores the ascending bond order -- which
//      makes the result independent of timing -- and shuffles it (Fisher-Yates, Lemire-bounded
//      Philox draws; the lanes of a warp walk the step index together so that all of them reach
//      a new group of four draws in the same iteration), and the CTA writes the super-bucket
//      out, coalesced.
// DRAM traffic: about 16 bytes per bond, all of it sequential.  M <= 2^24 bonds.
// ---------------------------------------------------------------------------
static constexpr int PH2_SUB_LOG_MAX = 7;
static constexpr int PH2_ID_BITS = 24;

struct Philox2Plan {
    int log2_buckets, sub_log, supers;
    int super_cap;
    size_t smem_bytes;
};

static Philox2Plan plan_philox2(int32_t M)
{
    Philox2Plan p{};
    int lb = 0;
    while (lb < 17 && ((long long)64 << lb) < M) ++lb;
    p.log2_buckets = lb;
    p.sub_log = lb < PH2_SUB_LOG_MAX ? lb : PH2_SUB_LOG_MAX;
    p.supers = 1 << (lb - p.sub_log);
    const double mean = (double)M / (double)p.supers;
    p.super_cap = p.supers == 1 ? std::max(M, 1) : (int)(mean + 10.0 * sqrt(mean) + 64.0);
    const size_t nsub = (size_t)1 << p.sub_log;
    p.smem_bytes = ((size_t)PH_WARPS + 1) * (p.supers + 1) * 4 + ((size_t)PH_WARPS + 1) * (nsub + 1) * 4 +
                   (size_t)p.super_cap * 4 + 64;
    return p;
}

__global__ void __launch_bounds__(PH_THREADS) perm_philox2_kernel(int32_t M, int32_t R, const uint32_t *seeds,
                                                                   int32_t *perms, int log2b, int sub_log,
                                                                   int super_cap)
{
    extern __shared__ __align__(16) uint32_t sm[];
    const int B = 1 << log2b;
    const uint32_t bmask = (uint32_t)B - 1u;
    const int NSUB = 1 << sub_log, SB = B >> sub_log;
    uint32_t *sstart = sm;                              // [SB + 1]  super-bucket starts
    uint32_t *wsuper = sstart + SB + 1;                 // [PH_WARPS][SB + 1]  counts, then the warps' append positions
    uint32_t *bstart = wsuper + PH_WARPS * (SB + 1);    // [NSUB + 1]  bucket starts inside the super-bucket
    uint32_t *wsub = bstart + NSUB + 1;                 // [PH_WARPS][NSUB + 1]
    uint32_t *buf = wsub + PH_WARPS * (NSUB + 1);       // [super_cap]
    __shared__ uint32_t scan_tot[PH_WARPS];

    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int per_warp = (((M + PH_WARPS - 1) / PH_WARPS) + 127) & ~127;
    const int w_lo = min(M, warp * per_warp), w_hi = min(M, w_lo + per_warp);

    // exclusive scan over `cnt` values spread one run of `per` per thread; returns this thread's
    // prefix (callers write the results)
    auto block_excl = [&](uint32_t mine) -> uint32_t {
        uint32_t incl = mine;
        for (int k = 1; k < 32; k <<= 1) {
            const uint32_t o = __shfl_up_sync(0xffffffffu, incl, k);
            if (lane >= k) incl += o;
        }
        __syncthreads();                        // (scan_tot of the previous use has been read)
        if (lane == 31) scan_tot[warp] = incl;
        __syncthreads();
        uint32_t pre = incl - mine;
        for (int w = 0; w < warp; ++w) pre += scan_tot[w];
        return pre;
    };

    for (int run = blockIdx.x; run < R; run += gridDim.x) {
        const uint32_t seed = seeds[run];
        int32_t *out = perms + (size_t)run * M;

        // ---- A: per-warp super-bucket counts -----------------------------------------------
        for (int i = t; i < (PH_WARPS + 1) * (SB + 1); i += PH_THREADS) sm[i] = 0;
        __syncthreads();
        for (int i0 = w_lo + 4 * lane; i0 < w_hi; i0 += 128) {
            uint32_t o[4];
            philox_buckets4(seed, (uint32_t)i0 >> 2, o);
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (i0 + q < w_hi) atomicAdd(&wsuper[warp * (SB + 1) + ((o[q] & bmask) >> sub_log)], 1u);
        }
        __syncthreads();
        {
            // super-bucket starts: thread t owns super-buckets [t*per, (t+1)*per); warp w appends to a
            // super-bucket behind the warps before it
            const int per = (SB + PH_THREADS - 1) / PH_THREADS;
            const int s_lo = min(SB, t * per), s_hi = min(SB, s_lo + per);
            uint32_t mine = 0;
            for (int sp = s_lo; sp < s_hi; ++sp)
                for (int w = 0; w < PH_WARPS; ++w) mine += wsuper[w * (SB + 1) + sp];
            uint32_t pos = block_excl(mine);
            for (int sp = s_lo; sp < s_hi; ++sp) {
                sstart[sp] = pos;
                for (int w = 0; w < PH_WARPS; ++w) {
                    const uint32_t c = wsuper[w * (SB + 1) + sp];
                    wsuper[w * (SB + 1) + sp] = pos;
                    pos += c;
                }
            }
            if (t == PH_THREADS - 1) sstart[SB] = (uint32_t)M;
        }
        __syncthreads();

        // ---- B: append every bond to its warp's segment of its super-bucket ------------------
        for (int blk = w_lo; blk < w_hi; blk += 128) {
            uint32_t o[4];
            philox_buckets4(seed, (uint32_t)(blk + 4 * lane) >> 2, o);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int i = blk + 4 * lane + q;
                const uint32_t b = o[q] & bmask;
                // (two bonds of one warp instruction that share a super-bucket may land in either
                // order: phase C sorts every bucket, so the result does not depend on it)
                if (i < w_hi)
                    out[atomicAdd(&wsuper[warp * (SB + 1) + (b >> sub_log)], 1u)] =
                        (int32_t)((uint32_t)i | ((b & (NSUB - 1)) << PH2_ID_BITS));
            }
        }
        __syncthreads();

        // ---- C: super-bucket by super-bucket ----------------------------------------------------
        for (int sp = 0; sp < SB; ++sp) {
            const uint32_t lo = sstart[sp], n = sstart[sp + 1] - lo;
            if (n > (uint32_t)super_cap) __trap();       // (ten standard deviations above the mean)
            // per-warp counts of the buckets: warp w owns the entries [w*ch, (w+1)*ch) of the super-bucket
            const uint32_t ch = ((n + PH_WARPS - 1) / PH_WARPS + 31u) & ~31u;
            const uint32_t c_lo = min(n, warp * ch), c_hi = min(n, c_lo + ch);
            for (int i = t; i < PH_WARPS * (NSUB + 1); i += PH_THREADS) wsub[i] = 0;
            __syncthreads();
            for (uint32_t k = c_lo + lane; k < c_hi; k += 32)
                atomicAdd(&wsub[warp * (NSUB + 1) + ((uint32_t)out[lo + k] >> PH2_ID_BITS)], 1u);
            __syncthreads();
            {
                uint32_t mine = 0;
                if (t < NSUB)
                    for (int w = 0; w < PH_WARPS; ++w) mine += wsub[w * (NSUB + 1) + t];
                uint32_t pos = block_excl(mine);
                if (t < NSUB) {
                    bstart[t] = pos;
                    for (int w = 0; w < PH_WARPS; ++w) {
                        const uint32_t c = wsub[w * (NSUB + 1) + t];
                        wsub[w * (NSUB + 1) + t] = pos;
                        pos += c;
                    }
                }
                if (t == 0) bstart[NSUB] = n;
            }
            __syncthreads();
            // placement: warps by their bases, the entries of one warp instruction in any order
            uint32_t v_next = c_lo + lane < c_hi ? (uint32_t)out[lo + c_lo + lane] : 0u;
            for (uint32_t k0 = c_lo; k0 < c_hi; k0 += 32) {
                const uint32_t k = k0 + lane;
                const bool ok = k < c_hi;
                const uint32_t v = v_next;
                if (k + 32 < c_hi) v_next = (uint32_t)out[lo + k + 32];     // the next step's load is in flight
                if (ok) buf[atomicAdd(&wsub[warp * (NSUB + 1) + (v >> PH2_ID_BITS)], 1u)] = v & ((1u << PH2_ID_BITS) - 1u);
            }
            __syncthreads();
            if (t < NSUB) {
                const int b = (sp << sub_log) + t;
                const uint32_t s0 = bstart[t], sz = bstart[t + 1] - s0;
                uint32_t *x = buf + s0;
                for (uint32_t k = 1; k < sz; ++k) {       // restore ascending bond order
                    const uint32_t v = x[k];
                    uint32_t j = k;
                    while (j > 0 && x[j - 1] > v) { x[j] = x[j - 1]; --j; }
                    x[j] = v;
                }
                const uint32_t kmax = __reduce_max_sync(__activemask(), sz);
                uint32_t o[4] = {0u, 0u, 0u, 0u};
                for (uint32_t k = kmax - 1; k >= 1 && kmax > 1; --k) {
                    if (((k & 3u) == 3u || k == kmax - 1) && (k & ~3u) < sz)
                        philox4x32_10(seed, PHILOX_KEY1, k >> 2, (uint32_t)b, 0u, 1u, o);
                    if (k < sz) {
                        const uint32_t w = (k & 2u) ? ((k & 1u) ? o[3] : o[2]) : ((k & 1u) ? o[1] : o[0]);
                        const uint32_t j = philox_bounded_w(seed, (uint32_t)b, k, w);
                        const uint32_t a = x[k], c = x[j];
                        x[k] = c; x[j] = a;
                    }
                }
            }
            __syncthreads();
            for (uint32_t k = t; k < n; k += PH_THREADS) out[lo + k] = (int32_t)buf[k];
            __syncthreads();
        }
    }
}

cudaError_t launch_perm_philox(int32_t M, int32_t R, const uint32_t *seeds, int32_t *perms,
                               cudaStream_t s, int *launches)
{
    *launches = 0;
    if (R <= 0 || M <= 0) return cudaSuccess;
    if ((long long)M > (1ll << PH2_ID_BITS)) return cudaErrorInvalidValue;      // (documented limit of this mode)
    const Philox2Plan p = plan_philox2(M);
    cudaError_t e = cudaFuncSetAttribute(perm_philox2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)p.smem_bytes);
    if (e != cudaSuccess) return e;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = std::min(R, sms * 8);
    perm_philox2_kernel<<<grid, PH_THREADS, p.smem_bytes, s>>>(M, R, seeds, perms, p.log2_buckets, p.sub_log,
                                                               p.super_cap);
    *launches = 1;
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// keyed bijection (Feistel network with cycle walking)
// ---------------------------------------------------------------------------
static constexpr int FEISTEL_ROUNDS = 20;
static constexpr int FE_THREADS = 256;
static constexpr int FE_ITEMS = 8;             // positions per thread

// (ROUNDS = FEISTEL_ROUNDS in production; the statistical tests also run deliberately weakened networks,
// PZ_FEISTEL_ROUNDS, to show that they notice)
template <int ROUNDS>
__global__ void __launch_bounds__(FE_THREADS) perm_feistel_kernel(int32_t M, int32_t R,
                                                                  const uint32_t *seeds, int32_t *perms,
                                                                  int k_bits)
{
    constexpr int FEISTEL_ROUNDS = ROUNDS;
    __shared__ uint32_t keys[FEISTEL_ROUNDS];
    const int run = blockIdx.y;
    if (threadIdx.x < (FEISTEL_ROUNDS + 3) / 4) {
        uint32_t o[4];
        philox4x32_10(seeds[run], PHILOX_KEY1, (uint32_t)threadIdx.x, 0u, 0u, 3u, o);
#pragma unroll
        for (int q = 0; q < 4; ++q) if (4 * threadIdx.x + q < FEISTEL_ROUNDS) keys[4 * threadIdx.x + q] = o[q];
    }
    __syncthreads();
    uint32_t key[FEISTEL_ROUNDS];
#pragma unroll
    for (int i = 0; i < FEISTEL_ROUNDS; ++i) key[i] = keys[i];

    const int a = k_bits >> 1, b = k_bits - a;
    const uint32_t mb = (1u << b) - 1u;            // right half
    const uint32_t mba = mb << a;                  // where the right half goes
    const uint32_t mk = k_bits >= 32 ? 0xffffffffu : (1u << k_bits) - 1u;
    const uint32_t two_a = 1u << a, two_32mb = 1u << (32 - b);
    int32_t *out = perms + (size_t)run * M;
    const int base = blockIdx.x * (FE_THREADS * FE_ITEMS) + threadIdx.x;
    // every lane walks its own list of positions: a point that lands at or above M
    // is walked on by that lane alone (no lane waits for another lane's walk)
    int n = base;
    const int n_end = min(M, base + FE_ITEMS * FE_THREADS);
    uint32_t x = (uint32_t)n;
    while (n < n_end) {
#pragma unroll
        for (int i = 0; i < FEISTEL_ROUNDS; ++i) {
            // bits above k carry garbage between rounds; no round reads them
            const uint64_t pr = (uint64_t)((x & mb) ^ key[i]) * 0xD2511F53u;
            const uint32_t y = x ^ (uint32_t)(pr >> 32) ^ (uint32_t)pr;
            // both shifts as multiplies (x << a = x * 2^a, y >> b = hi(y * 2^(32-b))): the
            // integer-multiply pipe is idle otherwise and the logic pipe is the bound
            uint32_t xa, s;
            asm("mul.lo.u32 %0, %1, %2;" : "=r"(xa) : "r"(x), "r"(two_a));
            asm("mul.hi.u32 %0, %1, %2;" : "=r"(s) : "r"(y), "r"(two_32mb));
            x = (xa & mba) | (s & ~mba);
        }
        x &= mk;
        if (x < (uint32_t)M) {
            out[n] = (int32_t)x;
            n += FE_THREADS;
            x = (uint32_t)n;
        }
    }
}

cudaError_t launch_perm_feistel(int32_t M, int32_t R, const uint32_t *seeds, int32_t *perms,
                                cudaStream_t s, int *launches)
{
    *launches = 0;
    if (R <= 0 || M <= 0) return cudaSuccess;
    int k = 2;
    while (((long long)1 << k) < M) ++k;
    const int per_block = FE_THREADS * FE_ITEMS;
    for (int r0 = 0; r0 < R; r0 += 65535) {         // gridDim.y limit
        const int rn = R - r0 < 65535 ? R - r0 : 65535;
        dim3 grid((unsigned)((M + per_block - 1) / per_block), (unsigned)rn);
        // (read per call: a test hook, not a tuning knob)
        const int rounds = getenv("PZ_FEISTEL_ROUNDS") ? atoi(getenv("PZ_FEISTEL_ROUNDS")) : FEISTEL_ROUNDS;
        switch (rounds) {
        case 2: perm_feistel_kernel<2><<<grid, FE_THREADS, 0, s>>>(M, rn, seeds + r0, perms + (size_t)r0 * M, k); break;
        case 3: perm_feistel_kernel<3><<<grid, FE_THREADS, 0, s>>>(M, rn, seeds + r0, perms + (size_t)r0 * M, k); break;
        case 4: perm_feistel_kernel<4><<<grid, FE_THREADS, 0, s>>>(M, rn, seeds + r0, perms + (size_t)r0 * M, k); break;
        case 6: perm_feistel_kernel<6><<<grid, FE_THREADS, 0, s>>>(M, rn, seeds + r0, perms + (size_t)r0 * M, k); break;
        case 8: perm_feistel_kernel<8><<<grid, FE_THREADS, 0, s>>>(M, rn, seeds + r0, perms + (size_t)r0 * M, k); break;
        default:
            perm_feistel_kernel<FEISTEL_ROUNDS><<<grid, FE_THREADS, 0, s>>>(M, rn, seeds + r0, perms + (size_t)r0 * M, k);
        }
        *launches += 1;
    }
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// Warp-cooperative Fisher-Yates: ONE WARP PER RUN, 32 shuffle steps at a time.
//
// The shuffle  for i = M-1 .. 1: j = draw(i); swap(x[i], x[j])  is a serial chain only through
// memory: the targets j come from the generator alone.  A warp therefore draws the targets of a
// ROW of up to 32 consecutive steps at once (lane l = the l-th step of the row, in stream order)
// and applies them together: the steps of a row commute unless two of them touch the same
// position, i.e. two steps share a target (found with match.any) or a step's target is the own
// position i of a later step of the row (a range test).  Both are rare while i is large (about
// 32^2 / i per row); a row with a collision is cut at the first step that collides with an
// earlier one of the same segment and applied segment by segment, in order -- so the result is
// exactly the sequential one for every M, also for the last few hundred steps, where nearly
// every row is cut.  Rows are applied one behind the draws: the targets of row g+1 are known
// (and their lines prefetched into L2) while row g waits for its loads, so a row costs about one
// L2 round trip for 32 steps instead of one dependent memory access per step and thread.
//
// Generators (template parameter):
//  * MT19937, NumPy's legacy stream bit for bit.  The 624-word state lives in REGISTERS, word
//    32 r + l in register r of lane l (20 registers; row 19 is half full).  A row of the next
//    generation needs words k, k+1 and k+397 (mod 624) only, i.e. the same row shifted by one
//    lane and one other row rotated by 13 (or 29) lanes: four shuffles for 32 words.  The
//    masked-rejection draw (numpy's legacy interval: smallest 2^k - 1 >= i, redraw above i)
//    couples a step to the number of draws accepted before it; inside a row that is a ballot
//    and a population count as long as no draw falls into the 31 values below the row's first i
//    (else, and whenever the mask changes inside the row, the row is decided lane by lane).
//  * Philox4x32-10 counter-based draws (PZ_PERM_PHILOX_FY): step i takes word i & 3 of counter
//    (i >> 2, 0, 0, 4) under key (seed, 'PERC'), unbiased by Lemire's method (rare rejections
//    take words of (i, attempt, 0, 5)); the tests hold a CPU restatement.
//
// No shared memory, 8 warps per CTA and at most 64 registers: one such CTA fits next to the
// sweep CTA that owns the rest of the SM, so the bond orders of the next batch of runs are
// generated underneath the sweep of the current one.
// ---------------------------------------------------------------------------
static constexpr int WP_WARPS = 8;             // warps per CTA of a shared launch
static constexpr int WP_WARPS_EXCL = 32;       // ... of a launch that has SMs to itself
static constexpr uint32_t FULL = 0xffffffffu;

struct WarpRow {
    uint32_t acc;                  // lanes that hold a step (warp-uniform)
    int32_t first;                 // i of the row's first step (warp-uniform)
    int32_t i, j;                  // this lane's step
};

// numpy's masked rejection for one row of 32-bit draws y (lane order = stream order)
__device__ __forceinline__ void wp_accept_mt(uint32_t y, int nvalid, int32_t &i, WarpRow &r, int lane)
{
    r.first = i;
    const uint32_t mask = 0xffffffffu >> __clz(i);
    const uint32_t m = y & mask;
    const bool le = lane < nvalid && m <= (uint32_t)i;
    const bool stable = i - 31 > (int32_t)(mask >> 1);       // one mask for the whole row, i stays >= 1
    const bool amb = le && (int32_t)m > i - 31;
    if (stable && !__any_sync(FULL, amb)) {
        r.acc = __ballot_sync(FULL, le);
        r.i = i - __popc(r.acc & ((1u << lane) - 1u));
        r.j = (int32_t)m;
        i -= __popc(r.acc);
        return;
    }
    uint32_t acc = 0;
    int32_t cur = i;
    r.i = 0; r.j = 0;
    for (int l = 0; l < nvalid; ++l) {
        const uint32_t yl = __shfl_sync(FULL, y, l);
        if (cur >= 1) {
            const uint32_t ml = yl & (0xffffffffu >> __clz(cur));
            if (ml <= (uint32_t)cur) {
                if (lane == l) { r.i = cur; r.j = (int32_t)ml; }
                acc |= 1u << l;
                --cur;
            }
        }
    }
    r.acc = acc;
    i = cur;
}

__device__ __forceinline__ uint32_t philox_fy_bounded(uint32_t seed, uint32_t i, uint32_t w)
{
    const uint32_t range = i + 1u;
    uint64_t m = (uint64_t)w * range;
    if ((uint32_t)m >= range) return (uint32_t)(m >> 32);      // cannot be below the threshold
    const uint32_t thresh = (0u - range) % range;
    if ((uint32_t)m >= thresh) return (uint32_t)(m >> 32);
    for (uint32_t attempt = 1;; ++attempt) {
        uint32_t r[4];
        philox4x32_10(seed, PHILOX_KEY1, i, attempt, 0u, 5u, r);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            m = (uint64_t)r[e] * range;
            if ((uint32_t)m >= thresh) return (uint32_t)(m >> 32);
        }
    }
}

// The shuffle of one run.  x = the warp's staging row (L2 resident, reused run after run): it holds
// the live part of the permutation, positions 0..i; position i is final after step i, so its value
// goes straight to the output row (a coalesced streaming store) and never back to the staging row.
// Rows are software pipelined: the loads of row g are in flight while row g+1 is drawn and checked
// for collisions; its stores follow, then the loads of row g+1 (behind the stores in program order
// and a __syncwarp: a later row may read what an earlier one wrote).
struct WarpShuffler {
    int32_t *x, *out;
    WarpRow p;                     // row whose loads are in flight (p.acc == 0: none)
    int32_t a, b;

    __device__ __forceinline__ void init(int32_t *x_, int32_t *out_) { x = x_; out = out_; p.acc = 0u; }
    __device__ __forceinline__ void finish(int lane) {
        if (!p.acc) return;
        if ((p.acc >> lane) & 1u) {
            __stcs(out + p.i, b);
            x[p.j] = a;
        }
        __syncwarp();
        p.acc = 0u;
    }
    __device__ __forceinline__ void push(const WarpRow &r, int lane) {
        if (!r.acc) { return; }
        const bool act = (r.acc >> lane) & 1u;
        const int32_t last = r.first - __popc(r.acc) + 1;
        const uint32_t lt = (1u << lane) - 1u;
        const uint32_t peers = __match_any_sync(FULL, act ? (uint32_t)r.j : (0x80000000u | (uint32_t)lane)) & r.acc;
        const bool dup = act && (peers & lt);                    // an earlier step has the same target
        const bool hit = act && r.j >= last && r.j < r.i;        // my target is a later step's own position
        const uint32_t confl = __ballot_sync(FULL, dup || hit);
        finish(lane);                                            // the previous row's stores
        if (!confl) {
            if (act) { a = x[r.i]; b = x[r.j]; }
            p = r;
            return;
        }
        // a row with a collision: cut at the first step that collides with an earlier one of the
        // same segment, segment by segment, in order
        uint32_t victim = 32u;                                   // the lane whose step sits on my target
        if (hit) victim = __fns(r.acc, 0, (r.first - r.j) + 1);
        int start = 0;
        while (start < 32) {
            const uint32_t seg = r.acc & ~((1u << start) - 1u);  // steps still to do
            if (!seg) break;
            const uint32_t c1 = __ballot_sync(FULL, act && lane >= start && (peers & lt & seg));
            int cut = c1 ? __ffs(c1) - 1 : 32;
            const uint32_t v = __reduce_min_sync(FULL, (hit && lane >= start) ? victim : 32u);
            if ((int)v < cut) cut = (int)v;
            if (act && lane >= start && lane < cut) {
                const int32_t va = x[r.i], vb = x[r.j];
                __stcs(out + r.i, vb);
                x[r.j] = va;
            }
            __syncwarp();
            start = cut;
        }
    }
    // after the last step position 0 is final as well
    __device__ __forceinline__ void done(int32_t M, int lane) {
        finish(lane);
        if (lane == 0 && M > 0) __stcs(out, x[0]);
        __syncwarp();
    }
};

__device__ __forceinline__ void wp_iota(int32_t *x, int32_t M, int lane)
{
    int4 *x4 = reinterpret_cast<int4 *>(x);                     // (staging rows are 128-byte aligned)
    const int n4 = M >> 2;
    for (int k = lane; k < n4; k += 32) x4[k] = make_int4(4 * k, 4 * k + 1, 4 * k + 2, 4 * k + 3);
    for (int k = (n4 << 2) + lane; k < M; k += 32) x[k] = k;
    __syncwarp();
}

// row r of the next twister generation (see the layout above); `s` is indexed statically
#define PZ_MT_ROW(r)                                                                                   \
    {                                                                                                  \
        const uint32_t cur = s[r];                                                                     \
        const uint32_t t1 = __shfl_down_sync(FULL, cur, 1);                                            \
        const uint32_t t2 = __shfl_sync(FULL, s[((r) + 1) % 20], 0);                                   \
        const uint32_t nxt = (lane == ((r) == 19 ? 15 : 31)) ? t2 : t1;                                \
        uint32_t far;                                                                                  \
        if ((r) <= 6) {                                                                                \
            const uint32_t fa = __shfl_sync(FULL, s[((r) + 12) % 20], (lane + 13) & 31);               \
            const uint32_t fb = __shfl_sync(FULL, s[((r) + 13) % 20], (lane + 13) & 31);               \
            far = lane < 19 ? fa : fb;                                                                 \
        } else if ((r) == 7) {                                                                         \
            const uint32_t fa = __shfl_sync(FULL, s[19], (lane + 13) & 31);                            \
            const uint32_t fb = __shfl_sync(FULL, s[0], (lane + 29) & 31);                             \
            far = lane < 3 ? fa : fb;                                                                  \
        } else {                                                                                       \
            const uint32_t fa = __shfl_sync(FULL, s[((r) + 12) % 20], (lane + 29) & 31);               \
            const uint32_t fb = __shfl_sync(FULL, s[((r) + 13) % 20], (lane + 29) & 31);               \
            far = lane < 3 ? fa : fb;      /* rows r - 8 and r - 7 (mod 20), already of this generation */ \
        }                                                                                              \
        const uint32_t yy = (cur & 0x80000000u) | (nxt & 0x7fffffffu);                                 \
        uint32_t v = far ^ (yy >> 1) ^ ((yy & 1u) ? 0x9908b0dfu : 0u);                                 \
        s[r] = v;                                                                                      \
        v ^= (v >> 11);                                                                                \
        v ^= (v << 7) & 0x9d2c5680u;                                                                   \
        v ^= (v << 15) & 0xefc60000u;                                                                  \
        v ^= (v >> 18);                                                                                \
        y[r] = v;                                                                                      \
    }

template <bool MT>
__global__ void __launch_bounds__(32 * WP_WARPS_EXCL, 1) perm_warp_kernel(int32_t M, int32_t R, const uint32_t *seeds,
                                                                      int32_t *perms, int32_t *stage,
                                                                      size_t stage_stride)
{
    const int lane = threadIdx.x & 31;
    const int wpc = blockDim.x >> 5;
    const int gw = blockIdx.x * wpc + (threadIdx.x >> 5), nw = gridDim.x * wpc;
    int32_t *x = stage + (size_t)gw * stage_stride;
    for (int run = gw; run < R; run += nw) {
        const uint32_t seed = seeds[run];
        wp_iota(x, M, lane);
        WarpShuffler S;
        S.init(x, perms + (size_t)run * M);
        int32_t i = M - 1;                   // next step to draw a target for
        if (MT) {
            uint32_t s[20], y[20];
            {
                // init_genrand: a serial recurrence, computed by every lane; lane l keeps word 32 r + l
                uint32_t v = seed;
#pragma unroll
                for (int r = 0; r < 20; ++r) {
                    s[r] = 0u;
#pragma unroll 1
                    for (int l = 0; l < 32; ++l) {
                        if (lane == l) s[r] = v;
                        v = 1812433253u * (v ^ (v >> 30)) + (uint32_t)(32 * r + l + 1);
                    }
                }
            }
            while (i >= 1) {
                PZ_MT_ROW(0) PZ_MT_ROW(1) PZ_MT_ROW(2) PZ_MT_ROW(3) PZ_MT_ROW(4)
                PZ_MT_ROW(5) PZ_MT_ROW(6) PZ_MT_ROW(7) PZ_MT_ROW(8) PZ_MT_ROW(9)
                PZ_MT_ROW(10) PZ_MT_ROW(11) PZ_MT_ROW(12) PZ_MT_ROW(13) PZ_MT_ROW(14)
                PZ_MT_ROW(15) PZ_MT_ROW(16) PZ_MT_ROW(17) PZ_MT_ROW(18) PZ_MT_ROW(19)
#pragma unroll 1
                for (int r = 0; r < 20 && i >= 1; ++r) {
                    WarpRow row;
                    wp_accept_mt(y[r], r == 19 ? 16 : 32, i, row, lane);
                    S.push(row, lane);
                }
            }
        } else {
            // lane L holds the Philox block of counter c0 - L: 128 steps (four rows) per evaluation
            uint32_t o[4] = {0u, 0u, 0u, 0u};
            int32_t c0 = -1;
            while (i >= 1) {
                WarpRow row;
                row.first = i;
                row.i = i - lane;
                const bool act = row.i >= 1;
                row.acc = __ballot_sync(FULL, act);
                const int32_t c_hi = i >> 2, c_lo = (i - 31 > 0 ? i - 31 : 0) >> 2;
                if (c0 < 0 || c0 < c_hi || c0 - c_lo > 31) {
                    c0 = c_hi;
                    if (c0 - lane >= 0) philox4x32_10(seed, PHILOX_KEY1, (uint32_t)(c0 - lane), 0u, 0u, 4u, o);
                }
                const int src = act ? c0 - (row.i >> 2) : 0;
                const uint32_t w0 = __shfl_sync(FULL, o[0], src), w1 = __shfl_sync(FULL, o[1], src);
                const uint32_t w2 = __shfl_sync(FULL, o[2], src), w3 = __shfl_sync(FULL, o[3], src);
                const uint32_t w = (row.i & 2) ? ((row.i & 1) ? w3 : w2) : ((row.i & 1) ? w1 : w0);
                row.j = act ? (int32_t)philox_fy_bounded(seed, (uint32_t)row.i, w) : 0;
                i -= __popc(row.acc);
                S.push(row, lane);
            }
        }
        S.done(M, lane);
    }
}

// Shape of a launch.  Shared (excl_sms == 0): PZ_WP_CTAS CTAs per SM (default 2; next to a sweep CTA
// one of them is resident at a time), PZ_WP_WARPS warps each (default 8).  Exclusive (excl_sms > 0):
// one CTA of 32 warps on each of excl_sms SMs, which it keeps to itself by asking for most of the
// SM's shared memory -- the sweep of the previous chunk runs on the other SMs.  Every warp owns one
// staging row.  Measured on one B200 (profiles/rng_r2.txt): the shuffle saturates at about 1200
// warps (2.2e10 bonds/s, bound by the random sector traffic of the staging rows); large graphs take
// fewer warps so that the staging rows stay below 2 GB.
static constexpr size_t WP_EXCL_SMEM = 160 * 1024;

void perm_warp_shape(int sms, int32_t M, int excl_sms, int *ctas, int *wpc, size_t *stride)
{
    static const int per_sm = getenv("PZ_WP_CTAS") ? std::max(1, atoi(getenv("PZ_WP_CTAS"))) : 2;
    static const int w = getenv("PZ_WP_WARPS") ? std::min(WP_WARPS, std::max(1, atoi(getenv("PZ_WP_WARPS")))) : 8;
    *ctas = excl_sms > 0 ? excl_sms : sms * per_sm;
    *wpc = excl_sms > 0 ? WP_WARPS_EXCL : w;
    *stride = ((size_t)std::max(M, 1) + 31) / 32 * 32;
    const size_t budget = (size_t)2 << 30;
    while (*wpc > 1 && (size_t)*ctas * *wpc * *stride * 4 > budget) *wpc >>= 1;
    while (excl_sms == 0 && *ctas > sms && (size_t)*ctas * *wpc * *stride * 4 > budget) *ctas -= sms;
    while (excl_sms == 0 && *ctas > 1 && (size_t)*ctas * *wpc * *stride * 4 > budget) *ctas >>= 1;
}

size_t perm_stage_ints(int sms, int32_t M, int excl_sms)
{
    int ctas, wpc; size_t stride;
    perm_warp_shape(sms, M, 0, &ctas, &wpc, &stride);
    size_t need = (size_t)ctas * wpc * stride;
    if (excl_sms > 0) {
        perm_warp_shape(sms, M, excl_sms, &ctas, &wpc, &stride);
        need = std::max(need, (size_t)ctas * wpc * stride);
    }
    return need;
}

static cudaError_t launch_perm_serial(int mode_mt, int32_t M, int32_t R, const uint32_t *seeds,
                                      int32_t *perms, int32_t *stage, int excl_sms, cudaStream_t s,
                                      int *launches)
{
    *launches = 0;
    if (R <= 0 || M <= 0) return cudaSuccess;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int ctas, wpc; size_t stride;
    perm_warp_shape(sms, M, excl_sms, &ctas, &wpc, &stride);
    const int need = (R + wpc - 1) / wpc;
    if (ctas > need) ctas = need;
    const size_t smem = excl_sms > 0 ? WP_EXCL_SMEM : 0;
    auto kern = mode_mt ? perm_warp_kernel<true> : perm_warp_kernel<false>;
    if (smem) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    kern<<<ctas, 32 * wpc, smem, s>>>(M, R, seeds, perms, stage, stride);
    *launches = 1;
    return cudaGetLastError();
}

cudaError_t launch_perm_mt19937(int32_t M, int32_t R, const uint32_t *seeds, int32_t *perms,
                                int32_t *stage, int excl_sms, cudaStream_t s, int *launches)
{
    return launch_perm_serial(1, M, R, seeds, perms, stage, excl_sms, s, launches);
}

cudaError_t launch_perm_philox_fy(int32_t M, int32_t R, const uint32_t *seeds, int32_t *perms,
                                  int32_t *stage, int excl_sms, cudaStream_t s, int *launches)
{
    return launch_perm_serial(0, M, R, seeds, perms, stage, excl_sms, s, launches);
}

// ---------------------------------------------------------------------------
// caller-supplied bond orders (PZ_PERM_HOST / PZ_PERM_DEVICE): every row must be a permutation
// of 0..M-1 -- an entry outside the range would index the bond list out of bounds in the sweep,
// a repeated one gives records that mean nothing.  flag bit 0: out of range, bit 1: repeated.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) validate_orders_kernel(int32_t M, int32_t R, const int32_t *perms,
                                                              uint32_t *bitmap, int words, int *flag)
{
    for (int run = blockIdx.y; run < R; run += gridDim.y) {
        const int32_t *row = perms + (size_t)run * M;
        uint32_t *bits = bitmap + (size_t)run * words;
        int bad = 0;
        for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < M; n += gridDim.x * blockDim.x) {
            const uint32_t e = (uint32_t)row[n];
            if (e >= (uint32_t)M) { bad |= 1; continue; }
            const uint32_t bit = 1u << (e & 31u);
            if (atomicOr(&bits[e >> 5], bit) & bit) bad |= 2;
        }
        if (bad) atomicOr(flag, bad);
    }
}

cudaError_t launch_validate_orders(int32_t M, int32_t R, const int32_t *perms, uint32_t *bitmap,
                                   int *flag, cudaStream_t s)
{
    if (R <= 0 || M <= 0) return cudaSuccess;
    const int words = (M + 31) / 32;
    cudaError_t e = cudaMemsetAsync(bitmap, 0, ((size_t)R * words + 1) * 4, s);
    if (e != cudaSuccess) return e;
    dim3 grid((unsigned)std::min(64, (M + 255) / 256), (unsigned)std::min(R, 16384));
    validate_orders_kernel<<<grid, 256, 0, s>>>(M, R, perms, bitmap, words, flag);
    return cudaGetLastError();
}

}  // namespace pz
