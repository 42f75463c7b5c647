// pz_rng.cu -- bond orders generated on the device (sm_100a).
//
// The reference draws the bond order of a run on the host,
// ``RandomState(seed).permutation(M)`` (percolate/hpc.py:195,206).  Two device
// replacements:
//
//  * perm_mt19937: NumPy's legacy stream reproduced bit for bit -- MT19937
//    seeded by init_genrand(seed), then the legacy shuffle (for i = M-1..1:
//    j = masked-rejection draw in [0, i]; swap).  One thread per run; the
//    twister state lives in thread-local memory, the permutation in the run's
//    row of the output (HBM-resident, latency hidden by running every run of
//    the batch concurrently).
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
//    Every Philox call therefore serves four bonds.
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
static constexpr int PH_THREADS = 256;
static constexpr int PH_WARPS = PH_THREADS / 32;

// the four bucket words of bonds 4g .. 4g+3
__device__ __forceinline__ void philox_buckets4(uint32_t seed, uint32_t g, uint32_t (&o)[4])
{
    philox4x32_10(seed, PHILOX_KEY1, g, 0u, 0u, 0u, o);
}

// unbiased j in [0, k] (Lemire).  `o` caches the four words of counter
// (k >> 2, bucket, 0, 1); `grp` is the group they belong to.
__device__ __forceinline__ uint32_t philox_bounded(uint32_t seed, uint32_t bucket, uint32_t k,
                                                   uint32_t (&o)[4], uint32_t &grp)
{
    const uint32_t range = k + 1u;
    if ((k >> 2) != grp) {
        grp = k >> 2;
        philox4x32_10(seed, PHILOX_KEY1, grp, bucket, 0u, 1u, o);
    }
    const uint32_t w = (k & 2u) ? ((k & 1u) ? o[3] : o[2]) : ((k & 1u) ? o[1] : o[0]);
    uint64_t m = (uint64_t)w * range;
    if ((uint32_t)m >= range) return (uint32_t)(m >> 32);      // cannot be below the threshold
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

struct PhiloxPlan {
    int log2_buckets;
    int cap;             // shared-memory capacity (entries) of one bucket in the shuffle phase
    int fy_threads;
    size_t smem_bytes;
};

static PhiloxPlan plan_philox(int32_t M)
{
    PhiloxPlan p{};
    int lb = 0;
    while (lb < 11 && ((long long)64 << lb) < M) ++lb;      // mean bucket ~64 until B = 2048
    p.log2_buckets = lb;
    const double mean = (double)M / (double)(1 << lb);
    int cap = (int)(mean + 8.0 * sqrt(mean) + 16.0);
    cap |= 1;                                               // odd stride: no systematic bank conflicts
    // the shuffle buffers reuse the space of the per-warp histograms, so that
    // three CTAs fit one SM (occupancy matters more than shuffle threads)
    const size_t hist = (size_t)PH_WARPS * ((size_t)1 << lb) * 4;
    size_t budget = hist > (size_t)32 * 1024 ? hist : (size_t)32 * 1024;
    int t = (int)(budget / ((size_t)cap * 4));
    if (t > PH_THREADS) t = PH_THREADS;
    if (t < 1) t = 1;
    p.cap = cap;
    p.fy_threads = t;
    const size_t fy = (size_t)t * cap * 4;
    p.smem_bytes = (hist > fy ? hist : fy) + ((size_t)(1 << lb) + 1) * 4 + 64;
    return p;
}

__global__ void __launch_bounds__(PH_THREADS) perm_philox_kernel(int32_t M, int32_t R,
                                                                  const uint32_t *seeds, int32_t *perms,
                                                                  int log2b, int cap, int fy_threads)
{
    extern __shared__ __align__(16) uint32_t sm[];
    const int B = 1 << log2b;
    const uint32_t bmask = (uint32_t)B - 1u;
    uint32_t *start = sm;                       // [B + 1]
    uint32_t *hist = sm + B + 1;                // [PH_WARPS][B] then per-warp bases
    uint32_t *fybuf = sm + B + 1;               // [fy_threads][cap]   (after the scatter)
    __shared__ uint32_t scan_tot[PH_WARPS];

    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    // contiguous element range of this warp (multiple of 128: one Philox call
    // per lane covers four consecutive bonds)
    const int per_warp = (((M + PH_WARPS - 1) / PH_WARPS) + 127) & ~127;
    const int w_lo = min(M, warp * per_warp), w_hi = min(M, w_lo + per_warp);

    for (int run = blockIdx.x; run < R; run += gridDim.x) {
        const uint32_t seed = seeds[run];
        int32_t *out = perms + (size_t)run * M;

        // ---- A: per-warp histograms of the bucket draws ----------------------
        for (int i = t; i < PH_WARPS * B; i += PH_THREADS) hist[i] = 0;
        __syncthreads();
        for (int i0 = w_lo + 4 * lane; i0 < w_hi; i0 += 128) {
            uint32_t o[4];
            philox_buckets4(seed, (uint32_t)i0 >> 2, o);
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (i0 + q < w_hi) atomicAdd(&hist[warp * B + (o[q] & bmask)], 1u);
        }
        __syncthreads();

        // ---- A': bucket starts (exclusive scan) and per-warp bases -----------
        {
            // thread t owns buckets [t*per, (t+1)*per)
            const int per = (B + PH_THREADS - 1) / PH_THREADS;
            const int b_lo = min(B, t * per), b_hi = min(B, b_lo + per);
            uint32_t mine = 0;
            for (int b = b_lo; b < b_hi; ++b)
                for (int w = 0; w < PH_WARPS; ++w) mine += hist[w * B + b];
            uint32_t incl = mine;
            for (int k = 1; k < 32; k <<= 1) {
                const uint32_t o = __shfl_up_sync(0xffffffffu, incl, k);
                if (lane >= k) incl += o;
            }
            if (lane == 31) scan_tot[warp] = incl;
            __syncthreads();
            uint32_t pre = incl - mine;
            for (int w = 0; w < warp; ++w) pre += scan_tot[w];
            for (int b = b_lo; b < b_hi; ++b) {
                start[b] = pre;
                uint32_t run_base = pre;
                for (int w = 0; w < PH_WARPS; ++w) {
                    const uint32_t h = hist[w * B + b];
                    hist[w * B + b] = run_base;
                    run_base += h;
                }
                pre = run_base;
            }
            if (t == PH_THREADS - 1) start[B] = (uint32_t)M;
        }
        __syncthreads();

        // ---- B: scatter.  Positions come from the per-warp bucket bases; two bonds
        // of one warp instruction that share a bucket may land in either order, which
        // phase C repairs by sorting the (almost sorted) bucket -- the order inside a
        // bucket is ascending bond index whatever the timing.
        for (int blk = w_lo; blk < w_hi; blk += 128) {
            uint32_t o[4];
            philox_buckets4(seed, (uint32_t)(blk + 4 * lane) >> 2, o);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int i = blk + 4 * lane + q;
                if (i < w_hi) out[atomicAdd(&hist[warp * B + (o[q] & bmask)], 1u)] = i;
            }
        }
        __syncthreads();

        // ---- C: Fisher-Yates inside every bucket, in shared memory -------------
        for (int b0 = 0; b0 < B; b0 += fy_threads) {
            const int b = b0 + t;
            if (t < fy_threads && b < B) {
                const uint32_t s0 = start[b], sz = start[b + 1] - s0;
                if (sz > 1) {
                    if ((int)sz <= cap) {
                        uint32_t *buf = fybuf + (size_t)t * cap;
                        for (uint32_t k = 0; k < sz; ++k) buf[k] = (uint32_t)out[s0 + k];
                        for (uint32_t k = 1; k < sz; ++k) {       // restore ascending bond order
                            const uint32_t v = buf[k];
                            uint32_t j = k;
                            while (j > 0 && buf[j - 1] > v) { buf[j] = buf[j - 1]; --j; }
                            buf[j] = v;
                        }
                        uint32_t o[4], grp = 0xffffffffu;
                        for (uint32_t k = sz - 1; k >= 1; --k) {
                            const uint32_t j = philox_bounded(seed, (uint32_t)b, k, o, grp);
                            const uint32_t a = buf[k], c = buf[j];
                            buf[k] = c; buf[j] = a;
                        }
                        for (uint32_t k = 0; k < sz; ++k) out[s0 + k] = (int32_t)buf[k];
                    } else {                      // over-full bucket: same sort + shuffle in place
                        for (uint32_t k = 1; k < sz; ++k) {
                            const int32_t v = out[s0 + k];
                            uint32_t j = k;
                            while (j > 0 && out[s0 + j - 1] > v) { out[s0 + j] = out[s0 + j - 1]; --j; }
                            out[s0 + j] = v;
                        }
                        uint32_t o[4], grp = 0xffffffffu;
                        for (uint32_t k = sz - 1; k >= 1; --k) {
                            const uint32_t j = philox_bounded(seed, (uint32_t)b, k, o, grp);
                            const int32_t a = out[s0 + k], c = out[s0 + j];
                            out[s0 + k] = c; out[s0 + j] = a;
                        }
                    }
                }
            }
        }
        __syncthreads();
    }
}

cudaError_t launch_perm_philox(int32_t M, int32_t R, const uint32_t *seeds, int32_t *perms,
                               cudaStream_t s, int *launches)
{
    *launches = 0;
    if (R <= 0 || M <= 0) return cudaSuccess;
    const PhiloxPlan p = plan_philox(M);
    cudaError_t e = cudaFuncSetAttribute(perm_philox_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)p.smem_bytes);
    if (e != cudaSuccess) return e;
    perm_philox_kernel<<<R, PH_THREADS, p.smem_bytes, s>>>(M, R, seeds, perms, p.log2_buckets, p.cap,
                                                           p.fy_threads);
    *launches = 1;
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// keyed bijection (Feistel network with cycle walking)
// ---------------------------------------------------------------------------
static constexpr int FEISTEL_ROUNDS = 20;
static constexpr int FE_THREADS = 256;
static constexpr int FE_ITEMS = 8;             // positions per thread

__global__ void __launch_bounds__(FE_THREADS) perm_feistel_kernel(int32_t M, int32_t R,
                                                                  const uint32_t *seeds, int32_t *perms,
                                                                  int k_bits)
{
    __shared__ uint32_t keys[FEISTEL_ROUNDS];
    const int run = blockIdx.y;
    if (threadIdx.x < FEISTEL_ROUNDS / 4) {
        uint32_t o[4];
        philox4x32_10(seeds[run], PHILOX_KEY1, (uint32_t)threadIdx.x, 0u, 0u, 3u, o);
#pragma unroll
        for (int q = 0; q < 4; ++q) keys[4 * threadIdx.x + q] = o[q];
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
        perm_feistel_kernel<<<grid, FE_THREADS, 0, s>>>(M, rn, seeds + r0, perms + (size_t)r0 * M, k);
        *launches += 1;
    }
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// NumPy legacy stream: MT19937 + masked-rejection Fisher-Yates
// ---------------------------------------------------------------------------
__global__ void iota_rows_kernel(int32_t M, int32_t R, int32_t *perms)
{
    const size_t total = (size_t)M * R;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x)
        perms[i] = (int32_t)(i % (size_t)M);
}

// One thread per run.  The shuffle is a serial chain, but only through memory: the
// target j of step i comes from the generator alone, so the targets of the next SWQ steps
// are drawn ahead of time and their x[j] loads are in flight while the current step
// swaps.  A step that writes a position a queued load has already read (its own target j,
// which now holds the old x[i]) forwards the new value to that queue entry, so the result
// is exactly the sequential one.
//
// The queue is a shift register of SWQ (target, prefetched value) pairs, head at index 0,
// empty entries marked by target -1.  push() executes the head's step (if any), shifts,
// and appends the new target at the tail; SWQ pushes of the empty marker drain it.
#ifndef PZ_MTQ
#define PZ_MTQ 8
#endif
static constexpr int SWQ = PZ_MTQ;

struct SwapQueue {
    int32_t jq[SWQ], vq[SWQ];
    int32_t *x;
    int32_t i;                           // step the head entry belongs to
    __device__ __forceinline__ void init(int32_t *x_, int32_t M) {
        x = x_; i = M - 1;
#pragma unroll
        for (int k = 0; k < SWQ; ++k) { jq[k] = -1; vq[k] = 0; }
    }
    // j < 0: the empty marker (drains the queue)
    __device__ __forceinline__ void push(int32_t jn) {
        const int32_t j = jq[0];
        int32_t a = 0;
        if (j >= 0) {                    // step i: swap x[i] and x[j]
            a = x[i];
            const int32_t b = (j == i) ? a : vq[0];
            x[i] = b;
            x[j] = a;
            --i;
            // x[i] runs down the row one element per step: fetch its line ahead of the dependent load
            if ((i & 7) == 7 && i >= 64)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(x + (i - 64)));
        }
        // shift, forwarding this step's write of x[j] to loads that came too early
        // (a queued target is below i, so it never equals position i)
#pragma unroll
        for (int k = 0; k + 1 < SWQ; ++k) {
            const int32_t jj = jq[k + 1];
            int32_t vv = vq[k + 1];
            if (jj == j) vv = a;
            jq[k] = jj; vq[k] = vv;
        }
        jq[SWQ - 1] = jn;
        vq[SWQ - 1] = jn >= 0 ? x[jn] : 0;     // behind this step's stores in program order
    }
    __device__ __forceinline__ void drain() {
#pragma unroll 1
        for (int k = 0; k < SWQ; ++k) push(-1);
    }
};

// NumPy's stream.  The twister state of a run is one CONTIGUOUS 2.5 KB row of global memory
// (thread-local arrays are interleaved word by word across the lanes of a warp, and lanes that
// have rejected different numbers of draws then touch 32 different lines per access).  The state
// is advanced in place EIGHT words at a time -- element k of the next generation needs only
// elements k, k+1 and k+397 (mod 624), so a block of eight is independent of itself -- instead
// of all 624 at once: with the block regeneration of the textbook code every lane of a warp
// reaches its regeneration at a different step (the lanes have rejected different numbers of
// draws), each of those runs with one active lane, and the warp pays 32 serial regenerations
// per 468 steps: that, not the shuffle, was 95 % of the kernel (8.6e9 bonds/s).
// Every lane then looks at the same draw q of its block in lock step; a draw is either
// accepted (masked value <= i: one shuffle step) or rejected (nothing happens), so there is
// no rejection LOOP for lanes to diverge in.
__global__ void __launch_bounds__(64) perm_mt19937_kernel(int32_t M, int32_t R, const uint32_t *seeds,
                                                           int32_t *perms, uint32_t *states)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
    for (int run = tid; run < R; run += nthr) {
        uint32_t *mt = states + (size_t)tid * 624;
        {
            uint32_t v = seeds[run];
            mt[0] = v;
            for (int i = 1; i < 624; ++i) { v = 1812433253u * (v ^ (v >> 30)) + (uint32_t)i; mt[i] = v; }
        }
        int k0 = 0;                          // next block of the state row (multiple of 8)
        SwapQueue q;
        q.init(perms + (size_t)run * M, M);
        int32_t i_fill = M - 1;              // next step to draw a target for
        while (i_fill >= 1) {
            uint32_t a[9], d[8];
            {
                const uint4 lo = *reinterpret_cast<const uint4 *>(mt + k0);
                const uint4 hi = *reinterpret_cast<const uint4 *>(mt + k0 + 4);
                a[0] = lo.x; a[1] = lo.y; a[2] = lo.z; a[3] = lo.w;
                a[4] = hi.x; a[5] = hi.y; a[6] = hi.z; a[7] = hi.w;
                a[8] = mt[k0 + 8 == 624 ? 0 : k0 + 8];
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                int km = k0 + e + 397;
                if (km >= 624) km -= 624;
                const uint32_t y = (a[e] & 0x80000000u) | (a[e + 1] & 0x7fffffffu);
                d[e] = mt[km] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
            *reinterpret_cast<uint4 *>(mt + k0) = make_uint4(d[0], d[1], d[2], d[3]);
            *reinterpret_cast<uint4 *>(mt + k0 + 4) = make_uint4(d[4], d[5], d[6], d[7]);
            k0 = k0 + 8 == 624 ? 0 : k0 + 8;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                uint32_t y = d[e];
                y ^= (y >> 11);
                y ^= (y << 7) & 0x9d2c5680u;
                y ^= (y << 15) & 0xefc60000u;
                y ^= (y >> 18);
                if (i_fill >= 1) {
                    // numpy legacy rk_interval: smallest 2^k - 1 >= i, masked rejection
                    const uint32_t j = y & (0xffffffffu >> __clz(i_fill));
                    if (j <= (uint32_t)i_fill) { q.push((int32_t)j); --i_fill; }
                }
            }
        }
        q.drain();
    }
}

// Textbook Fisher-Yates with counter-based draws (PZ_PERM_PHILOX_FY): for i = M-1 .. 1,
// j = unbiased (Lemire) draw in [0, i] from word i & 3 of Philox4x32-10 counter (i >> 2, 0, 0, 4)
// under key (seed, 'PERC'); on the (rare, < 2^-15) Lemire rejection words 0.. of counters
// (i, attempt >= 1, 0, 5).  One thread per run, no shared memory and no state: like the
// MT19937 kernel it runs underneath the sweep of the previous batch of runs.
// oracle/pz_oracle.c restates it on the CPU (philox_fy_permutation).
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

__global__ void __launch_bounds__(64) perm_philox_fy_kernel(int32_t M, int32_t R, const uint32_t *seeds,
                                                             int32_t *perms)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
    for (int run = tid; run < R; run += nthr) {
        const uint32_t seed = seeds[run];
        SwapQueue q;
        q.init(perms + (size_t)run * M, M);
        for (int32_t g = (M - 1) >> 2; g >= 0; --g) {
            uint32_t o[4];
            philox4x32_10(seed, PHILOX_KEY1, (uint32_t)g, 0u, 0u, 4u, o);
#pragma unroll
            for (int e = 3; e >= 0; --e) {
                const int32_t i = 4 * g + e;
                if (i >= 1 && i <= M - 1) q.push((int32_t)philox_fy_bounded(seed, (uint32_t)i, o[e]));
            }
        }
        q.drain();
    }
}

// runs of one launch of the thread-per-run kernels: at most SERIAL_CTAS_PER_SM CTAs of 64 threads
// per SM, so that a sweep CTA (46 K registers, all of the shared memory) still fits next to them
static constexpr int SERIAL_CTAS_PER_SM = 4;
int perm_serial_capacity(int sms) { return sms * SERIAL_CTAS_PER_SM * 64; }

static cudaError_t launch_perm_serial(int mode_mt, int32_t M, int32_t R, const uint32_t *seeds,
                                      int32_t *perms, cudaStream_t s, int *launches)
{
    *launches = 0;
    if (R <= 0 || M <= 0) return cudaSuccess;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int ctas = (R + 63) / 64;
    // more runs than one co-resident launch holds: threads take several runs in turn
    // (PZ_SERIAL_FULL=1: as many threads as runs -- when nothing has to fit next to the kernel)
    static const bool full = getenv("PZ_SERIAL_FULL") && atoi(getenv("PZ_SERIAL_FULL"));
    if (!full && ctas > sms * SERIAL_CTAS_PER_SM) ctas = sms * SERIAL_CTAS_PER_SM;
    uint32_t *states = nullptr;
    cudaError_t e = cudaSuccess;
    if (mode_mt) {
        e = cudaMallocAsync(&states, (size_t)ctas * 64 * 624 * sizeof(uint32_t), s);
        if (e != cudaSuccess) return e;
    }
    iota_rows_kernel<<<1184, 256, 0, s>>>(M, R, perms);
    if (mode_mt) perm_mt19937_kernel<<<ctas, 64, 0, s>>>(M, R, seeds, perms, states);
    else perm_philox_fy_kernel<<<ctas, 64, 0, s>>>(M, R, seeds, perms);
    *launches = 2;
    e = cudaGetLastError();
    if (mode_mt) {
        const cudaError_t e2 = cudaFreeAsync(states, s);
        if (e == cudaSuccess) e = e2;
    }
    return e;
}

cudaError_t launch_perm_mt19937(int32_t M, int32_t R, const uint32_t *seeds, int32_t *perms,
                                cudaStream_t s, int *launches)
{
    return launch_perm_serial(1, M, R, seeds, perms, s, launches);
}

cudaError_t launch_perm_philox_fy(int32_t M, int32_t R, const uint32_t *seeds, int32_t *perms,
                                  cudaStream_t s, int *launches)
{
    return launch_perm_serial(0, M, R, seeds, perms, s, launches);
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
