// pz_common.cuh -- shared definitions of the sm_100a Newman-Ziff kernels.
//
// Vocabulary (the reference's, percolate/hpc.py:194-307): a RUN adds the M
// BONDS of the graph one at a time in a permuted order; after each bond the
// run reports the largest-cluster size, the spanning flag and the k = 0..4
// moments of the cluster-size distribution without one largest cluster.
//
// The sweep kernel does not write those statistics.  It writes one MERGE
// RECORD per bond: 0 if the bond joined nothing, else the sizes (minus one) of
// the two clusters it joined.  Everything the reference reports is a prefix
// scan of that stream (see pz_stats.cu):
//     c[n]   = number of merges among bonds 1..n
//     max[n] = max(1, max over merges <= n of (w0 + w1))
//     S_k[n] = N + sum over merges <= n of ((w0+w1)^k - w0^k - w1^k)   (mod 2^64)
//     moments[k][n] = S_k[n] - max[n]^k                                 (mod 2^64)
// which equals the reference's incremental bookkeeping (hpc.py:283-305)
// because that bookkeeping maintains "sum over all clusters but one largest".
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace pz {

// ---- merge records ---------------------------------------------------------
// rec32 (N <= 65536): bit31 valid | bits 16..30 min(size)-1 | bits 0..15 max(size)-1
// rec64 (any N)     : bit63 valid | bits 32..62 min(size)-1 | bits 0..31 max(size)-1
__host__ __device__ __forceinline__ uint32_t rec32_pack(uint32_t a_m1, uint32_t b_m1) {
    uint32_t lo = a_m1 < b_m1 ? a_m1 : b_m1, hi = a_m1 < b_m1 ? b_m1 : a_m1;
    return 0x80000000u | (lo << 16) | hi;
}
__host__ __device__ __forceinline__ uint64_t rec64_pack(uint32_t a_m1, uint32_t b_m1) {
    uint32_t lo = a_m1 < b_m1 ? a_m1 : b_m1, hi = a_m1 < b_m1 ? b_m1 : a_m1;
    return 0x8000000000000000ull | ((uint64_t)lo << 32) | hi;
}
template <class RecT> struct RecCodec;
template <> struct RecCodec<uint32_t> {
    __host__ __device__ static __forceinline__ bool valid(uint32_t r) { return r >> 31; }
    __host__ __device__ static __forceinline__ uint32_t w_small(uint32_t r) { return ((r >> 16) & 0x7fffu) + 1; }
    __host__ __device__ static __forceinline__ uint32_t w_large(uint32_t r) { return (r & 0xffffu) + 1; }
};
template <> struct RecCodec<uint64_t> {
    __host__ __device__ static __forceinline__ bool valid(uint64_t r) { return r >> 63; }
    __host__ __device__ static __forceinline__ uint32_t w_small(uint64_t r) { return (uint32_t)((r >> 32) & 0x7fffffffu) + 1; }
    __host__ __device__ static __forceinline__ uint32_t w_large(uint64_t r) { return (uint32_t)(r & 0xffffffffu) + 1; }
};

static constexpr uint32_t NSPAN_NEVER = 0xffffffffu;

// running statistics of one run along n (all wrap mod 2^64 like the
// reference's uint64 arithmetic, hpc.py:286,297,302)
struct RunState {
    uint32_t c;      // merges so far
    uint32_t mx;     // largest cluster
    uint64_t s2, s3, s4;
    __host__ __device__ void init(uint32_t N) { c = 0; mx = 1; s2 = s3 = s4 = N; }
    __host__ __device__ __forceinline__ void merge(uint32_t wa, uint32_t wb) {
        // with w = a + b and t = a b:  w^2 - a^2 - b^2 = 2t,  w^3 - a^3 - b^3 = 3tw,
        // w^4 - a^4 - b^4 = 2t (2 w^2 - t)   (identities in Z, hence mod 2^64)
        const uint32_t w = wa + wb;
        const uint64_t t = (uint64_t)wa * wb, w2 = (uint64_t)w * w;
        s2 += 2 * t;
        s3 += 3 * t * w;
        s4 += 2 * t * (2 * w2 - t);
        c += 1;
        if (w > mx) mx = w;
    }
    // moments[k], k = 0..4 (hpc.py:214 at n = 0, :283-305 afterwards)
    __host__ __device__ __forceinline__ void moments(uint32_t N, uint64_t m[5]) const {
        uint64_t x = mx, x2 = x * x;
        m[0] = (uint64_t)N - 1 - c;
        m[1] = (uint64_t)N - x;
        m[2] = s2 - x2;
        m[3] = s3 - x2 * x;
        m[4] = s4 - x2 * x2;
    }
};

}  // namespace pz
