// pz_sweep.cu -- the per-run union-find sweep (sm_100a).
//
// Replaces the hot loop of the reference, percolate/hpc.py:249-307 (twin:
// percolate/percolate.py:298-356): one Python iteration per bond over a dict
// union-find.  Here one CTA (or, in the A/B baseline kernel, one warp) owns one
// run and adds bonds a batch at a time:
//
//   1. every lane looks up the endpoints of its bond and finds both roots in
//      parallel (weighted quick-union, path halving).  Concurrent halving
//      writes are benign: they only ever replace a parent by an ancestor.
//   2. lanes whose roots differ are merge candidates.  A candidate may merge
//      immediately iff no LOWER lane touches either of its roots: then every
//      earlier bond of the batch works on disjoint clusters, so the sizes it
//      sees are exactly the ones the sequential reference sees.  This is
//      decided with one atomicMin claim per root in a hashed shared-memory
//      table (hash collisions only delay a lane, never break the order).
//   3. what is left (chains through a common cluster, typically the giant
//      one) is replayed in bond order, the whole warp walking together.
//
// The parent/size array of the run lives in shared memory whenever it fits
// (uint16 entries + one flag byte per node; N <= 65536 covers the L = 256 square
// lattice in 192 KB),
// otherwise in a per-warp slab of global memory that stays L2-resident.
// Output per bond is one merge record (see pz_common.cuh); per run one word,
// the first n at which the two spanning sides are joined (hpc.py:269-274).
#include <cstdio>
#include <cstdlib>
#include "pz_common.cuh"
#include "pz_internal.h"

namespace pz {

static constexpr uint32_t CLAIM_FREE = 0xffffffffu;
// claim keys = (epoch << 10) | position in the batch; the epoch counts down one per round and is
// rebased (claim table cleared) at a batch boundary once it drops below EPOCH_LOW
static constexpr uint32_t EPOCH_LOW = 0x00001000u;      // (first epoch: SweepArgs::epoch_start)
#ifndef PZ_STRICT_FENCE
#define PZ_STRICT_FENCE 0
#endif

__device__ __forceinline__ uint32_t align16(uint32_t x) { return (x + 15u) & ~15u; }
static inline size_t align16h(size_t x) { return (x + 15) & ~(size_t)15; }

// ---------------------------------------------------------------------------
// Stores.  Common surface:
//   init(lane)                 every node its own cluster of size 1
//   find(x, tok) -> root       tok = opaque root token (carries size-1)
//   size_m1(tok)
//   unite(ra, ta, rb, tb, track) -> side mask of the merged cluster (0 if !track)
// ---------------------------------------------------------------------------
struct StoreS16 {
    using Rec = uint32_t;
    using Edge = uint32_t;
    uint16_t *val;        // bit 15: root; low 15 bits: size-1 (root) or parent
    uint32_t *sides2;     // 2-bit side fields
    int32_t N;
    const uint32_t *sides_init;

    static size_t slice_bytes(int32_t N) {
        return align16h((size_t)N * 2) + align16h((size_t)((N + 15) / 16) * 4);
    }
    __device__ void bind(unsigned char *slice, const SweepArgs &a, int) {
        N = a.N;
        val = reinterpret_cast<uint16_t *>(slice);
        sides2 = reinterpret_cast<uint32_t *>(slice + align16((uint32_t)N * 2));
        sides_init = a.sides2;
    }
    __device__ void init(int tid, int nthr) {
        uint32_t *v32 = reinterpret_cast<uint32_t *>(val);
        for (int i = tid; i < (N + 1) / 2; i += nthr) v32[i] = 0x80008000u;
        if (sides_init)
            for (int i = tid; i < (N + 15) / 16; i += nthr) sides2[i] = sides_init[i];
    }
    __device__ __forceinline__ void find_pair(uint32_t &x, uint32_t &y, uint32_t &tx, uint32_t &ty) const {
        x = find(x, tx);
        y = find(y, ty);
    }
    __device__ __forceinline__ uint32_t find(uint32_t x, uint32_t &tok) const {
        uint32_t vx = val[x];
        while (!(vx & 0x8000u)) {
            const uint32_t p = vx, vp = val[p];
            if (vp & 0x8000u) { x = p; vx = vp; break; }
            val[x] = (uint16_t)vp;          // halve: parent[x] = grandparent
            x = vp;
            vx = val[x];
        }
        tok = vx;
        return x;
    }
    // walk two representatives up to their roots in lock step (no halving)
    __device__ __forceinline__ void find2(uint32_t &x, uint32_t &y, uint32_t &tx, uint32_t &ty) const {
        uint32_t vx = val[x], vy = val[y];
        while (!(vx & vy & 0x8000u)) {
            if (!(vx & 0x8000u)) x = vx;
            if (!(vy & 0x8000u)) y = vy;
            vx = val[x];
            vy = val[y];
        }
        tx = vx; ty = vy;
    }
    static __device__ __forceinline__ uint32_t size_m1(uint32_t tok) { return tok & 0x7fffu; }
    __device__ __forceinline__ uint32_t side_of(uint32_t r) const {
        return (sides2[r >> 4] >> ((r & 15u) * 2)) & 3u;
    }
    __device__ __forceinline__ uint32_t sides_of_root(uint32_t r, uint32_t) const { return side_of(r); }
    __device__ __forceinline__ void make_child(uint32_t x, uint32_t parent) { val[x] = (uint16_t)parent; }
    __device__ __forceinline__ void set_root(uint32_t r, uint32_t sz_m1, uint32_t add_sides) {
        val[r] = (uint16_t)(0x8000u | sz_m1);
        if (add_sides) atomicOr(&sides2[r >> 4], add_sides << ((r & 15u) * 2));
    }
    __device__ __forceinline__ uint32_t unite(uint32_t ra, uint32_t ta, uint32_t rb, uint32_t tb,
                                              bool track) {
        const uint32_t sa = size_m1(ta), sb = size_m1(tb);
        const uint32_t big = sa >= sb ? ra : rb, small = sa >= sb ? rb : ra;
        val[small] = (uint16_t)big;
        val[big] = (uint16_t)(0x8000u | (sa + sb + 1));
        uint32_t m = 0;
        if (track) {
            const uint32_t mb = side_of(big);
            m = mb | side_of(small);
            if (m != mb) atomicOr(&sides2[big >> 4], m << ((big & 15u) * 2));
        }
        return m;
    }
};

struct StoreS16B {
    using Rec = uint32_t;
    using Edge = uint32_t;
    uint16_t *val;        // size-1 (root) or parent
    uint8_t *flag;        // bit 0: root; bits 1-2: spanning sides touched by the cluster (roots)
    int32_t N;
    const uint32_t *sides_init;

    // token of a root: flag << 16 | size-1
    static size_t slice_bytes(int32_t N) { return align16h((size_t)N * 2) + align16h((size_t)N); }
    __device__ void bind(unsigned char *slice, const SweepArgs &a, int) {
        N = a.N;
        val = reinterpret_cast<uint16_t *>(slice);
        flag = slice + align16((uint32_t)N * 2);
        sides_init = a.sides2;
    }
    __device__ void init(int tid, int nthr) {
        uint4 *v128 = reinterpret_cast<uint4 *>(val);
        const int n128 = (N * 2 + 15) / 16;
        for (int i = tid; i < n128; i += nthr) v128[i] = make_uint4(0, 0, 0, 0);
        uint32_t *f32 = reinterpret_cast<uint32_t *>(flag);
        for (int i = tid; i < (N + 3) / 4; i += nthr) {      // four nodes per word
            uint32_t w = 0x01010101u;
            if (sides_init) {
                const uint32_t s = sides_init[i >> 2] >> ((i & 3) * 8);      // 4 x 2 bits
                w |= ((s & 3u) << 1) | (((s >> 2) & 3u) << 9) | (((s >> 4) & 3u) << 17) |
                     (((s >> 6) & 3u) << 25);
            }
            f32[i] = w;
        }
    }
    __device__ __forceinline__ void find_pair(uint32_t &x, uint32_t &y, uint32_t &tx, uint32_t &ty) const {
        x = find(x, tx);
        y = find(y, ty);
    }
    __device__ __forceinline__ uint32_t find(uint32_t x, uint32_t &tok) const {
        uint32_t fx = flag[x], vx = val[x];
        while (!(fx & 1u)) {
            const uint32_t p = vx;
            const uint32_t fp = flag[p], vp = val[p];
            if (fp & 1u) { x = p; fx = fp; vx = vp; break; }
            val[x] = (uint16_t)vp;              // halve: parent[x] = grandparent
            x = vp;
            fx = flag[x];
            vx = val[x];
        }
        tok = (fx << 16) | vx;
        return x;
    }
    __device__ __forceinline__ void find2(uint32_t &x, uint32_t &y, uint32_t &tx, uint32_t &ty) const {
        uint32_t fx = flag[x], fy = flag[y], vx = val[x], vy = val[y];
        while (!(fx & fy & 1u)) {
            if (!(fx & 1u)) x = vx;
            if (!(fy & 1u)) y = vy;
            fx = flag[x]; fy = flag[y];
            vx = val[x]; vy = val[y];
        }
        tx = (fx << 16) | vx;
        ty = (fy << 16) | vy;
    }
    static __device__ __forceinline__ uint32_t size_m1(uint32_t tok) { return tok & 0xffffu; }
    __device__ __forceinline__ uint32_t sides_of_root(uint32_t, uint32_t tok) const { return (tok >> 17) & 3u; }
    __device__ __forceinline__ void make_child(uint32_t x, uint32_t parent) {
        val[x] = (uint16_t)parent;
        flag[x] = 0;
    }
    __device__ __forceinline__ void set_root(uint32_t r, uint32_t sz_m1, uint32_t add_sides) {
        // only the designated thread of a star round calls this
        val[r] = (uint16_t)sz_m1;
        if (add_sides) flag[r] = (uint8_t)(flag[r] | (add_sides << 1));
    }
    // the winner owns both roots: plain stores, no atomics
    __device__ __forceinline__ uint32_t unite(uint32_t ra, uint32_t ta, uint32_t rb, uint32_t tb,
                                              bool) {
        const uint32_t sa = size_m1(ta), sb = size_m1(tb);
        const uint32_t big = sa >= sb ? ra : rb, small = sa >= sb ? rb : ra;
        const uint32_t m = ((ta | tb) >> 17) & 3u;
        val[small] = (uint16_t)big;
        flag[small] = 0;
        val[big] = (uint16_t)(sa + sb + 1);
        if (m != (((sa >= sb ? ta : tb) >> 17) & 3u)) flag[big] = (uint8_t)(1u | (m << 1));
        return m;
    }
};

struct StoreG32 {
    using Rec = uint64_t;
    using Edge = uint2;
    uint32_t *val;        // root: bit31 | sides << 29 | size-1 ; else parent
    int32_t N;
    const uint32_t *sides_init;

    static size_t slice_bytes(int32_t) { return 0; }
    __device__ void bind(unsigned char *, const SweepArgs &a, int gwarp) {
        N = a.N;
        val = a.gscratch + (size_t)gwarp * (size_t)a.N;
        sides_init = a.sides2;
    }
    __device__ void init(int tid, int nthr) {
        for (int i = tid; i < N; i += nthr) {
            uint32_t s = sides_init ? (sides_init[i >> 4] >> ((i & 15) * 2)) & 3u : 0u;
            val[i] = 0x80000000u | (s << 29);
        }
    }
    __device__ __forceinline__ uint32_t find(uint32_t x, uint32_t &tok) const {
        uint32_t vx = val[x];
        while (!(vx >> 31)) {
            const uint32_t p = vx, vp = val[p];
            if (vp >> 31) { x = p; vx = vp; break; }
            val[x] = vp;
            x = vp;
            vx = val[x];
        }
        tok = vx;
        return x;
    }
    // both endpoints of a bond at once: two independent chains in lock step (the
    // store is L2 resident, so the two dependent-load chains overlap), path
    // splitting: every node visited is re-pointed at its grandparent
    __device__ __forceinline__ void find_pair(uint32_t &x, uint32_t &y, uint32_t &tx, uint32_t &ty) const {
        uint32_t px = 0xffffffffu, py = 0xffffffffu;
        uint32_t vx = val[x], vy = val[y];
        while (!((vx & vy) >> 31)) {
            if (!(vx >> 31)) { if (px != 0xffffffffu) val[px] = vx; px = x; x = vx; }
            if (!(vy >> 31)) { if (py != 0xffffffffu) val[py] = vy; py = y; y = vy; }
            vx = val[x];
            vy = val[y];
        }
        tx = vx; ty = vy;
    }
    __device__ __forceinline__ void find2(uint32_t &x, uint32_t &y, uint32_t &tx, uint32_t &ty) const {
        uint32_t vx = val[x], vy = val[y];
        while (!((vx & vy) >> 31)) {
            if (!(vx >> 31)) x = vx;
            if (!(vy >> 31)) y = vy;
            vx = val[x];
            vy = val[y];
        }
        tx = vx; ty = vy;
    }
    static __device__ __forceinline__ uint32_t size_m1(uint32_t tok) { return tok & 0x1fffffffu; }
    __device__ __forceinline__ uint32_t sides_of_root(uint32_t, uint32_t tok) const { return (tok >> 29) & 3u; }
    __device__ __forceinline__ void make_child(uint32_t x, uint32_t parent) { val[x] = parent; }
    __device__ __forceinline__ void set_root(uint32_t r, uint32_t sz_m1, uint32_t add_sides) {
        // only the designated thread of a star round calls this: plain read-modify-write
        const uint32_t old = val[r];
        val[r] = 0x80000000u | (old & 0x60000000u) | (add_sides << 29) | sz_m1;
    }
    __device__ __forceinline__ uint32_t unite(uint32_t ra, uint32_t ta, uint32_t rb, uint32_t tb,
                                              bool) {
        const uint32_t sa = size_m1(ta), sb = size_m1(tb);
        const uint32_t big = sa >= sb ? ra : rb, small = sa >= sb ? rb : ra;
        const uint32_t m = ((ta | tb) >> 29) & 3u;
        val[small] = big;
        val[big] = 0x80000000u | (m << 29) | (sa + sb + 1);
        return m;
    }
};

template <class Rec> __device__ __forceinline__ Rec make_rec(uint32_t a, uint32_t b);
template <> __device__ __forceinline__ uint32_t make_rec<uint32_t>(uint32_t a, uint32_t b) { return rec32_pack(a, b); }
template <> __device__ __forceinline__ uint64_t make_rec<uint64_t>(uint32_t a, uint32_t b) { return rec64_pack(a, b); }

__device__ __forceinline__ void edge_uv(uint32_t e, uint32_t &u, uint32_t &v) { u = e & 0xffffu; v = e >> 16; }
__device__ __forceinline__ void edge_uv(uint2 e, uint32_t &u, uint32_t &v) { u = e.x; v = e.y; }
__device__ __forceinline__ uint32_t claim_slot(uint32_t r, int log2) {
    return (r * 0x9E3779B1u) >> (32 - log2);
}

// ---------------------------------------------------------------------------
// merge phase of one batch (warp-wide).  On entry every valid lane holds
// representatives (ancestors-or-self) of its bond's two endpoints.  Rounds:
//   every pending lane walks its representatives up to the current roots;
//   lanes whose roots differ are candidates; a candidate may merge now iff no
//   LOWER pending candidate touches either of its roots (then all earlier
//   bonds of the batch work on disjoint clusters, so the sizes it records are
//   the ones the sequential reference sees); the rest retries next round.
// The lowest pending candidate always merges, so the loop terminates.  With
// one or two candidates the decision is made with shuffles, otherwise with
// one atomicMin claim per root in a hashed shared-memory table (a hash
// collision only delays a lane, it never breaks the order).
// Returns the lane's merge record; span_n = first row at which the two
// spanning sides are joined by a merge of this batch (or NSPAN_NEVER).
// ---------------------------------------------------------------------------
template <class Store>
__device__ __forceinline__ typename Store::Rec
merge_batch(Store &st, uint32_t *claim, int clog, int lane, bool valid, int n, uint32_t ru,
            uint32_t rv, bool track, int any3, uint32_t &span_n)
{
    using Rec = typename Store::Rec;
    Rec rec = 0;
    span_n = NSPAN_NEVER;
    bool pending = valid;
    for (;;) {
        uint32_t tu = 0, tv = 0;
        if (pending) {
            st.find2(ru, rv, tu, tv);
            pending = ru != rv;
        }
        const uint32_t cmask = __ballot_sync(0xffffffffu, pending);
        if (!cmask) break;
        const int ncand = __popc(cmask);
        bool win;
        if (ncand <= 2) {
            const int l0 = __ffs(cmask) - 1;
            const uint32_t a0 = __shfl_sync(0xffffffffu, ru, l0), b0 = __shfl_sync(0xffffffffu, rv, l0);
            win = pending && (lane == l0 || (ru != a0 && ru != b0 && rv != a0 && rv != b0));
        } else {
            uint32_t su = 0, sv = 0;
            if (pending) {
                su = claim_slot(ru, clog);
                sv = claim_slot(rv, clog);
                atomicMin(&claim[su], (uint32_t)lane);
                atomicMin(&claim[sv], (uint32_t)lane);
            }
            __syncwarp();
            win = pending && claim[su] == (uint32_t)lane && claim[sv] == (uint32_t)lane;
            __syncwarp();
            if (pending) { claim[su] = CLAIM_FREE; claim[sv] = CLAIM_FREE; }
        }
        if (win) {
            rec = make_rec<Rec>(Store::size_m1(tu), Store::size_m1(tv));
            const uint32_t m = st.unite(ru, tu, rv, tv, track);
            if (track && (m == 3u || any3)) span_n = (uint32_t)n + 1;
            pending = false;
        }
        // every candidate merged: done (also orders the winners' writes
        // before the next round's reads)
        if (__ballot_sync(0xffffffffu, win) == cmask) break;
    }
    if (track) span_n = __reduce_min_sync(0xffffffffu, span_n);
    return rec;
}

// ---------------------------------------------------------------------------
// single-warp kernel: one warp = one run (kept for small graphs and as the
// A/B baseline of the team kernel below)
// ---------------------------------------------------------------------------
template <class Store>
__global__ void sweep_kernel(SweepArgs a, uint32_t slice_bytes)
{
    extern __shared__ __align__(16) unsigned char smem[];
    using Rec = typename Store::Rec;
    using Edge = typename Store::Edge;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wpc = blockDim.x >> 5;
    const int gwarp = blockIdx.x * wpc + warp, nwarps = gridDim.x * wpc;
    const int M = a.M;
    const int clog = a.claim_log2;

    unsigned char *slice = smem + (size_t)warp * slice_bytes;
    uint32_t *claim = reinterpret_cast<uint32_t *>(slice);
    Store st;
    st.bind(slice + (sizeof(uint32_t) << clog), a, gwarp);
    for (int i = lane; i < (1 << clog); i += 32) claim[i] = CLAIM_FREE;

    const Edge *edges = reinterpret_cast<const Edge *>(a.edges);
    const bool spanning = a.sides2 != nullptr;

    for (int run = gwarp; run < a.R; run += nwarps) {
        st.init(lane, 32);
        __syncwarp();
        const int32_t *perm = a.perms + (size_t)run * M;
        Rec *rec_out = reinterpret_cast<Rec *>(a.recs) + (size_t)run * M;
        uint32_t nspan = NSPAN_NEVER;
        bool track = spanning;                 // warp-uniform

        // two-deep software pipeline: perm[n] -> edges[perm[n]] -> use
        Edge uv_next = Edge();
        int32_t e_next = 0;
        if (lane < M) uv_next = __ldg(&edges[__ldcs(&perm[lane])]);
        if (lane + 32 < M) e_next = __ldcs(&perm[lane + 32]);

        for (int n0 = 0; n0 < M; n0 += 32) {
            const int n = n0 + lane;           // bond index; row index is n + 1
            const bool valid = n < M;
            const Edge uv = uv_next;
            if (n + 32 < M) uv_next = __ldg(&edges[e_next]);
            if (n + 64 < M) e_next = __ldcs(&perm[n + 64]);

            uint32_t ru = 0, rv = 0, t;
            if (valid) {
                uint32_t u, v;
                edge_uv(uv, u, v);
                ru = st.find(u, t);
                rv = st.find(v, t);
            }
            uint32_t span_n;
            const Rec rec = merge_batch(st, claim, clog, lane, valid, n, ru, rv, track, a.any3, span_n);
            if (track && span_n != NSPAN_NEVER) { nspan = span_n; track = false; }
            if (valid) __stcs(&rec_out[n], rec);
        }
        if (lane == 0) a.nspan[run] = nspan;
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------
// CTA kernel: one CTA = one run, every thread one bond of a batch of
// CTA_THREADS bonds, all warps in lock step:
//
//   find    every thread resolves both endpoints of its bond (path halving;
//           no merges are in flight, so concurrent halving is benign)
//   round   pending threads (roots differ) post an atomicMin claim, keyed by
//           their position in the batch, on the hashed slot of each root;
//           barrier; a thread whose two slots carry its own key merges -- no
//           earlier bond of the batch touches its clusters, so the sizes it
//           records are the sequential ones; barrier; losers walk up to the new
//           roots and go again.  The earliest pending bond always wins.
//   star    bonds that join a cluster to the current largest cluster (the
//           HUB) would serialise, one per round, once the giant cluster exists.
//           They claim only their other root; if every earlier pending bond
//           of the batch has merged or merges in this round, they all merge
//           at once and a block-wide ordered prefix sum of the attached sizes
//           gives each of them the hub size the sequential order would see.
//
// Claim keys carry a decreasing epoch in their upper bits, so claims of earlier
// rounds never need to be cleared.  The four warp schedulers of the SM work on
// the same run, and every dependent-latency chain (find, claim, merge) is paid
// once per CTA_THREADS bonds instead of once per 32.
// ---------------------------------------------------------------------------
static constexpr int CTA_MAX_WARPS = 32;

struct CtaShared {
    unsigned long long hub_key;               // (size << 32) | root of the largest cluster seen
    unsigned long long scan_tot[CTA_MAX_WARPS];
    uint32_t span_min;
    uint32_t bmin;                            // lowest blocked position of the round
    uint32_t star_epoch;                      // epoch of the last round in which a star bond was pending
};

template <class Store, int CTA_WARPS>
__global__ void __launch_bounds__(32 * CTA_WARPS, 1) sweep_cta_kernel(SweepArgs a, uint32_t store_bytes)
{
    constexpr int CTA_THREADS = 32 * CTA_WARPS;
    extern __shared__ __align__(16) unsigned char smem[];
    using Rec = typename Store::Rec;
    using Edge = typename Store::Edge;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int M = a.M;
    const int clog = a.claim_log2;

    uint32_t *claim = reinterpret_cast<uint32_t *>(smem);
    unsigned char *p = smem + (sizeof(uint32_t) << clog);
    Store st;
    st.bind(p, a, blockIdx.x);
    CtaShared *sh = reinterpret_cast<CtaShared *>(p + store_bytes);

    const Edge *edges = reinterpret_cast<const Edge *>(a.edges);
    const bool spanning = a.sides2 != nullptr;

    for (int run = blockIdx.x; run < a.R; run += gridDim.x) {
        st.init(tid, CTA_THREADS);
        for (int i = tid; i < (1 << clog); i += CTA_THREADS) claim[i] = CLAIM_FREE;
        if (tid == 0) {
            sh->span_min = NSPAN_NEVER;
            sh->bmin = 0xffffffffu;
            sh->star_epoch = 0xffffffffu;
            sh->hub_key = 1ull << 32;          // node 0, size 1
        }
        __syncthreads();
        const int32_t *perm = a.perms + (size_t)run * M;
        Rec *rec_out = reinterpret_cast<Rec *>(a.recs) + (size_t)run * M;
        bool track = spanning;                 // CTA-uniform
        uint32_t epoch = a.epoch_start;        // decreasing: newer claims always win over stale ones
#ifdef PZ_TIMING
        long long tm_find = 0, tm_init = clock64(), tm_rounds = 0, tm_t = 0, tm_seg[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tm_p = 0, tm_star = 0, tm_late = 0, tm_nlate = 0, tm_b0 = 0; int tm_pc = 0;
#define TM(k) { const long long _c = clock64(); tm_seg[k] += _c - tm_p; tm_p = _c; }
#else
#define TM(k)
#endif

        // two-deep software pipeline: perm[n] -> edges[perm[n]] -> use
        Edge uv_next = Edge();
        int32_t e_next = 0;
        if (tid < M) uv_next = __ldg(&edges[__ldcs(&perm[tid])]);
        if (tid + CTA_THREADS < M) e_next = __ldcs(&perm[tid + CTA_THREADS]);

        for (int n0 = 0; n0 < M; n0 += CTA_THREADS) {
            const int n = n0 + tid;            // bond index; row index is n + 1
            const bool valid = n < M;
            const Edge uv = uv_next;
            if (n + CTA_THREADS < M) uv_next = __ldg(&edges[e_next]);
            if (n + 2 * CTA_THREADS < M) e_next = __ldcs(&perm[n + 2 * CTA_THREADS]);
            // a batch takes at most CTA_THREADS rounds: rebase the claim epoch long before it
            // can run out (graphs with millions of merging rounds per run)
            if (epoch < EPOCH_LOW) {
                __syncthreads();
                for (int i = tid; i < (1 << clog); i += CTA_THREADS) claim[i] = CLAIM_FREE;
                if (tid == 0) sh->star_epoch = 0xffffffffu;
                epoch = a.epoch_start;
                __syncthreads();
            }

            uint32_t ru = 0, rv = 0, tu = 0, tv = 0;
#ifdef PZ_TIMING
            tm_t = clock64(); tm_b0 = tm_t;
#endif
            if (valid) {
                edge_uv(uv, ru, rv);
                st.find_pair(ru, rv, tu, tv);
            }
            bool pending = valid && ru != rv;
            Rec rec = 0;
#ifdef PZ_TIMING
            tm_find += clock64() - tm_t;
            tm_p = clock64();
            tm_pc = __syncthreads_count(pending);
            TM(7)
#endif
            for (;;) {
#ifdef PZ_TIMING
                ++tm_rounds; tm_p = clock64();
#endif
                // a warp without pending bonds only takes part in the barriers
                const bool warp_has = __any_sync(0xffffffffu, pending);
                const uint32_t key = (epoch << 10) | (uint32_t)tid;
                uint32_t hub = 0, o = 0, to = 0, th = 0, su = 0, sv = 0;
                bool star = false;
                if (warp_has) {
                    hub = (uint32_t)sh->hub_key;
                    // star bond: one side is the hub; o = the other root
                    star = pending && (ru == hub || rv == hub);
                    o = ru == hub ? rv : ru; to = ru == hub ? tv : tu;
                    th = ru == hub ? tu : tv;
                    if (pending) {
                        su = claim_slot(star ? o : ru, clog);
                        sv = claim_slot(star ? o : rv, clog);
                        atomicMin(&claim[su], key);
                        if (!star) atomicMin(&claim[sv], key);
                        else sh->star_epoch = epoch;            // this round needs the star barrier
                    }
                }
                TM(0)
                if (!__syncthreads_or(pending)) break;          // nothing (left) to merge
                TM(1)
                const bool star_round = sh->star_epoch == epoch;
                bool own = false;
                if (warp_has) {
                    own = pending && claim[su] == key && claim[sv] == key;
                    if (star_round && pending && !own) atomicMin(&sh->bmin, (uint32_t)tid);
                }
                TM(2)
                // winners that are not star bonds own both roots: they merge right away, no
                // earlier bond of the batch touches their clusters (star bonds never do either:
                // those touch the hub and a root they own themselves)
                bool won = false;
                if (own && !star) {
                    rec = make_rec<Rec>(Store::size_m1(tu), Store::size_m1(tv));
                    const uint32_t sz = Store::size_m1(tu) + Store::size_m1(tv) + 2;
                    const uint32_t m = st.unite(ru, tu, rv, tv, track);
                    if (track && (m == 3u || a.any3)) atomicMin(&sh->span_min, (uint32_t)n + 1);
                    const uint32_t big = Store::size_m1(tu) >= Store::size_m1(tv) ? ru : rv;
                    const unsigned long long hk = ((unsigned long long)sz << 32) | big;
                    if (hk > sh->hub_key) atomicMax(&sh->hub_key, hk);
                    won = true;
                }
                // rounds without star bonds skip this barrier
                int nstar = 0;
                if (star_round) nstar = __syncthreads_count(own && star);
                TM(3)
                TM(4)
#ifdef PZ_TIMING
                if (nstar) ++tm_star;
#endif
                if (nstar) {
                    // star bonds merge together iff no earlier bond of the batch is blocked
                    const bool sw = own && star && (uint32_t)tid < sh->bmin;
                    unsigned long long v = 0ull, incl = 0ull;
                    // the hub's side bits as of the start of the round: read BEFORE the barrier
                    // below, behind which the first star bond publishes the whole round's bits
                    // (StoreS16 keeps them in a live shared-memory word, not in the token)
                    uint32_t hs = 0u;
                    if (__any_sync(0xffffffffu, sw)) {
                        const uint32_t so = sw ? st.sides_of_root(o, to) : 0u;
                        if (sw && track) hs = st.sides_of_root(hub, th);
                        v = sw ? ((unsigned long long)(Store::size_m1(to) + 1) |
                                  ((unsigned long long)(so & 1u) << 40) |
                                  ((unsigned long long)(so >> 1) << 50)) : 0ull;
                        incl = v;
#pragma unroll
                        for (int k = 1; k < 32; k <<= 1) {
                            const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, k);
                            if (lane >= k) incl += t;
                        }
                    }
                    if (lane == 31) sh->scan_tot[warp] = incl;
                    __syncthreads();
                    if (sw) {
                        unsigned long long pre = incl - v, total = 0;
#pragma unroll
                        for (int w = 0; w < CTA_WARPS; ++w) {
                            const unsigned long long t = sh->scan_tot[w];
                            if (w < warp) pre += t;
                            total += t;
                        }
                        const uint32_t hub_m1 = Store::size_m1(th);     // hub size - 1 at round start
                        const uint32_t pre_sz = (uint32_t)(pre & 0xffffffffffull);
                        rec = make_rec<Rec>(Store::size_m1(to), hub_m1 + pre_sz);
                        st.make_child(o, hub);
                        if (track) {
                            const unsigned long long in = pre + v;
                            const uint32_t m = hs | (((in >> 40) & 0x3ffu) ? 1u : 0u) |
                                               (((in >> 50) & 0x3ffu) ? 2u : 0u);
                            if (m == 3u || a.any3) atomicMin(&sh->span_min, (uint32_t)n + 1);
                        }
                        if (pre_sz == 0) {      // first star bond of the round: publish the hub
                            const uint32_t tot_sz = (uint32_t)(total & 0xffffffffffull);
                            const uint32_t add = (((total >> 40) & 0x3ffu) ? 1u : 0u) |
                                                 (((total >> 50) & 0x3ffu) ? 2u : 0u);
                            st.set_root(hub, hub_m1 + tot_sz, track ? add : 0u);
                            // the star phase has one writer of the hub: plain compare and store, no CAS loop
                            const unsigned long long nk = ((unsigned long long)(hub_m1 + tot_sz + 1) << 32) | hub;
                            if (nk > sh->hub_key) sh->hub_key = nk;
                        }
                        won = true;
                    }
                }
                TM(5)
                if (won) pending = false;
                --epoch;
                if (tid == 0) sh->bmin = 0xffffffffu;           // read only between the two barriers above
                if (!__syncthreads_or(pending)) break;          // every candidate merged
                TM(6)
                if (pending) {                                  // walk up to the new roots
                    st.find2(ru, rv, tu, tv);
                    pending = ru != rv;
                }
                TM(7)
            }
#ifdef PZ_TIMING
            if (tm_pc < CTA_THREADS / 5) { tm_late += clock64() - tm_b0; ++tm_nlate; }
#endif
            if (track && sh->span_min != NSPAN_NEVER) track = false;
            if (valid) __stcs(&rec_out[n], rec);
        }
        __syncthreads();
#ifdef PZ_TIMING
        if ((tid & 127) == 0 && blockIdx.x == 0 && run < (int)gridDim.x)
            printf("warp %2d: total %lld find %lld (%lld iterations, %lld star) claim %lld bar1 %lld own %lld bar2 %lld merge %lld star %lld bar3 %lld find2+findbar %lld\n",
                   warp, clock64() - tm_init, tm_find, tm_rounds, tm_star,
                   tm_seg[0], tm_seg[1], tm_seg[2], tm_seg[3], tm_seg[4], tm_seg[5], tm_seg[6], tm_seg[7]),
            printf("         batches with < T/5 pending bonds: %lld, %lld cycles\n", tm_nlate, tm_late);
#endif
        if (tid == 0) a.nspan[run] = sh->span_min;
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------
// CTA kernel with FINDER warps (the shape used when one run fills the shared memory of an SM).
// The lock-step rounds leave three quarters of the issue slots idle, and the finds of a batch
// (a fifth of the lock-step time) depend on nothing but the forest.  So 8 extra warps walk the
// endpoints of batch b+1 up the forest WHILE the 16 main warps run the rounds of batch b and hand
// the nodes they reach -- representatives: ancestors-or-self, valid whatever merges happen in
// between -- to the main warps through a 2 KB buffer.  The main warps start a batch by walking
// those representatives up to the current roots (0-1 steps) and never execute the find phase.
//
// Finders read the forest while merges are in flight.  A node turns from root into child by two
// stores, parent first, then the root flag with RELEASE order; a finder reads the flag with ACQUIRE
// order before the value, so a cleared flag always comes with the parent pointer.  The other
// direction (flag still set, value already a parent) makes the finder stop at a node that was a
// root a moment ago, which is a valid representative.  Path halving is done by the finders only
// and touches non-root entries, which merges never write.
//
// Named barriers: 1 = the 512 main threads (rounds), 2 = representatives of a batch are in the
// buffer (finders arrive, mains wait), 3 = the buffer has been read (mains arrive, finders wait).
// ---------------------------------------------------------------------------
#ifndef PZ_FW_WARPS
#define PZ_FW_WARPS 4
#endif
#ifndef PZ_FW_MAIN_WARPS
#define PZ_FW_MAIN_WARPS 16
#endif
static constexpr int FW_MAIN_WARPS = PZ_FW_MAIN_WARPS, FW_FIND_WARPS = PZ_FW_WARPS;
static constexpr int FW_MAIN = 32 * FW_MAIN_WARPS, FW_FIND = 32 * FW_FIND_WARPS, FW_ALL = FW_MAIN + FW_FIND;
static constexpr int FW_BPT = FW_MAIN / FW_FIND;          // bonds per finder thread and batch

__device__ __forceinline__ void nb_sync(int id, int n) {
    asm volatile("barrier.cta.sync %0, %1;" ::"r"(id), "r"(n) : "memory");
}
__device__ __forceinline__ void nb_arrive(int id, int n) {
    asm volatile("barrier.cta.arrive %0, %1;" ::"r"(id), "r"(n) : "memory");
}
__device__ __forceinline__ bool nb_or(int id, int n, bool pred) {
    uint32_t r;
    asm volatile("{\n\t.reg .pred q, o;\n\tsetp.ne.u32 q, %3, 0;\n\tbarrier.cta.red.or.pred o, %1, %2, q;\n\t"
                 "selp.u32 %0, 1, 0, o;\n\t}"
                 : "=r"(r) : "r"(id), "r"(n), "r"((uint32_t)pred) : "memory");
    return r != 0;
}
__device__ __forceinline__ int nb_count(int id, int n, bool pred) {
    uint32_t r;
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %3, 0;\n\tbarrier.cta.red.popc.u32 %0, %1, %2, q;\n\t}"
                 : "=r"(r) : "r"(id), "r"(n), "r"((uint32_t)pred) : "memory");
    return (int)r;
}

// StoreS16B whose root -> child transitions are release stores, plus the finders' acquire walk
struct StoreS16BF : StoreS16B {
    __device__ __forceinline__ void clear_flag_release(uint32_t x) {
        const uint32_t addr = (uint32_t)__cvta_generic_to_shared(flag + x);
        asm volatile("st.release.cta.shared.u8 [%0], %1;" ::"r"(addr), "r"(0u) : "memory");
    }
    __device__ __forceinline__ uint32_t flag_acquire(uint32_t x) const {
        uint32_t f;
        const uint32_t addr = (uint32_t)__cvta_generic_to_shared(flag + x);
        asm volatile("ld.acquire.cta.shared.u8 %0, [%1];" : "=r"(f) : "r"(addr) : "memory");
        return f;
    }
    __device__ __forceinline__ void make_child(uint32_t x, uint32_t parent) {
        val[x] = (uint16_t)parent;
        clear_flag_release(x);
    }
    __device__ __forceinline__ uint32_t unite(uint32_t ra, uint32_t ta, uint32_t rb, uint32_t tb, bool) {
        const uint32_t sa = size_m1(ta), sb = size_m1(tb);
        const uint32_t big = sa >= sb ? ra : rb, small = sa >= sb ? rb : ra;
        const uint32_t m = ((ta | tb) >> 17) & 3u;
        val[small] = (uint16_t)big;
        clear_flag_release(small);
        val[big] = (uint16_t)(sa + sb + 1);
        if (m != (((sa >= sb ? ta : tb) >> 17) & 3u)) flag[big] = (uint8_t)(1u | (m << 1));
        return m;
    }
    // K walks of find_rep at once: the dependent shared-memory loads of the K chains are issued
    // back to back, so a step costs one load latency for all of them instead of K
    template <int K>
    __device__ __forceinline__ void find_rep_multi(uint32_t (&x)[K]) {
        uint32_t f[K];
        bool any = false;
#pragma unroll
        for (int k = 0; k < K; ++k) { f[k] = flag_acquire(x[k]); any |= !(f[k] & 1u); }
        while (any) {
            uint32_t p[K], fp[K], g[K];
#pragma unroll
            for (int k = 0; k < K; ++k) p[k] = (f[k] & 1u) ? x[k] : (uint32_t)val[x[k]];   // parent (flag was clear)
#pragma unroll
            for (int k = 0; k < K; ++k) fp[k] = (f[k] & 1u) ? 1u : flag_acquire(p[k]);
#pragma unroll
            for (int k = 0; k < K; ++k) g[k] = (fp[k] & 1u) ? p[k] : (uint32_t)val[p[k]];  // grandparent
            any = false;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                if (!(f[k] & 1u)) {
                    if (fp[k] & 1u) { x[k] = p[k]; f[k] = 1u; }          // the parent is a root: done
                    else { val[x[k]] = (uint16_t)g[k]; x[k] = g[k]; f[k] = 0u; }   // halve, go on from g
                }
            }
#pragma unroll
            for (int k = 0; k < K; ++k)
                if (!(f[k] & 1u)) { f[k] = flag_acquire(x[k]); any |= !(f[k] & 1u); }
        }
    }
    // finder: walk x up while merges may be in flight; returns an ancestor-or-self of x that
    // was a root when its flag was read (path halving on the way)
    __device__ __forceinline__ uint32_t find_rep(uint32_t x) {
        uint32_t fx = flag_acquire(x);
        while (!(fx & 1u)) {
            const uint32_t p = val[x];                  // a parent pointer: the flag was clear
            const uint32_t fp = flag_acquire(p);
            if (fp & 1u) { x = p; break; }
            const uint32_t g = val[p];
            val[x] = (uint16_t)g;                       // halve: parent[x] = grandparent
            x = g;
            fx = flag_acquire(x);
        }
        return x;
    }
};

// (launch bound = the CTA plus one 256-thread CTA of the bond-order kernel: that caps the registers
// at 72 per thread, so that perm_feistel_kernel of the next chunk still finds room on the SM)
// 1: a finder thread walks the two endpoints of a bond up the forest together
#ifndef PZ_FIND_ILP
#define PZ_FIND_ILP 1
#endif
#ifndef PZ_TAIL_MAX
#define PZ_TAIL_MAX 32                        // pending bonds at or below which warp 0 finishes the batch
#endif
// hand-over area of the tail mode (see below): at most 32 pending bonds of a batch, in batch order
struct TailShared {
    uint32_t wcnt[FW_MAIN_WARPS];             // pending bonds per main warp
    uint32_t iu[32], iv[32], in[32];          // representatives and bond index of item k
    unsigned long long irec[32];              // its merge record (0 = joined nothing)
    uint32_t epoch;                           // epoch after the tail rounds
};
static_assert(sizeof(TailShared) <= 768, "TailShared must fit the planner's reserve");

template <class Store>
__global__ void __launch_bounds__(FW_ALL + 256, 1) sweep_fw_kernel(SweepArgs a, uint32_t store_bytes)
{
    extern __shared__ __align__(16) unsigned char smem[];
    using Rec = typename Store::Rec;
    using Edge = typename Store::Edge;
    constexpr int CTA_WARPS = FW_MAIN_WARPS;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool finder = tid >= FW_MAIN;
    const int M = a.M;
    const int clog = a.claim_log2;

    // the last FW_MAIN words of the claim table's 2^clog are the hand-over buffer (a slot index
    // beyond the shortened table is folded back), so the kernel needs no more shared memory than
    // the plain one and the bond-order kernel of the next chunk still fits next to it
    uint32_t *claim = reinterpret_cast<uint32_t *>(smem);
    const uint32_t claim_n = (1u << clog) - FW_MAIN;
    uint32_t *reps = claim + claim_n;                                          // [FW_MAIN] ru | rv << 16
    unsigned char *p = smem + (sizeof(uint32_t) << clog);
    Store st;
    st.bind(p, a, blockIdx.x);
    CtaShared *sh = reinterpret_cast<CtaShared *>(p + store_bytes);
    TailShared *ts = reinterpret_cast<TailShared *>(p + store_bytes + 320);

    const Edge *edges = reinterpret_cast<const Edge *>(a.edges);
    const bool spanning = a.sides2 != nullptr;
    const int nb = (M + FW_MAIN - 1) / FW_MAIN;

    for (int run = blockIdx.x; run < a.R; run += gridDim.x) {
        st.init(tid, FW_ALL);
        for (int i = tid; i < (int)claim_n; i += FW_ALL) claim[i] = CLAIM_FREE;
        if (tid == 0) {
            sh->span_min = NSPAN_NEVER;
            sh->bmin = 0xffffffffu;
            sh->star_epoch = 0xffffffffu;
            sh->hub_key = 1ull << 32;          // node 0, size 1
        }
        __syncthreads();
        const int32_t *perm = a.perms + (size_t)run * M;

        if (finder) {
            // ---- finder warps: FW_BPT bonds per thread and batch -----------------------------
            const int ft = tid - FW_MAIN;
            Edge uv_next[FW_BPT];
            int32_t e_next[FW_BPT];
#pragma unroll
            for (int q = 0; q < FW_BPT; ++q) {
                uv_next[q] = Edge(); e_next[q] = 0;
                const int n = ft + q * FW_FIND;
                if (n < M) uv_next[q] = __ldg(&edges[__ldcs(&perm[n])]);
                if (n + FW_MAIN < M) e_next[q] = __ldcs(&perm[n + FW_MAIN]);
            }
            for (int b = 0; b < nb; ++b) {
                uint32_t out[FW_BPT];
#pragma unroll
                for (int q = 0; q < FW_BPT; ++q) {
                    const int n = b * FW_MAIN + ft + q * FW_FIND;
                    const Edge uv = uv_next[q];
                    if (n + FW_MAIN < M) uv_next[q] = __ldg(&edges[e_next[q]]);
                    if (n + 2 * FW_MAIN < M) e_next[q] = __ldcs(&perm[n + 2 * FW_MAIN]);
                    uint32_t u = 0, v = 0;
                    if (n < M) {
                        edge_uv(uv, u, v);
#if PZ_FIND_ILP
                        uint32_t x[2] = {u, v};
                        st.template find_rep_multi<2>(x);
                        u = x[0]; v = x[1];
#else
                        u = st.find_rep(u);
                        v = st.find_rep(v);
#endif
                    }
                    out[q] = u | (v << 16);
                }
                nb_sync(3, FW_ALL);                     // the buffer of the previous batch has been read
#pragma unroll
                for (int q = 0; q < FW_BPT; ++q) reps[ft + q * FW_FIND] = out[q];
                nb_arrive(2, FW_ALL);                   // representatives of batch b are in the buffer
            }
        } else {
            // ---- main warps: rounds -----------------------------------------------------------
            Rec *rec_out = reinterpret_cast<Rec *>(a.recs) + (size_t)run * M;
            bool track = spanning;                 // uniform over the main warps
            uint32_t epoch = a.epoch_start;        // decreasing: newer claims always win over stale ones
            if (nb > 0) nb_arrive(3, FW_ALL);      // the buffer is free
#ifdef PZ_TIMING
            long long fw_wait = 0, fw_walk = 0, fw_tail = 0, fw_ntail = 0, fw_trounds = 0, fw_cta = 0, fw_ncta = 0, fw_titems = 0;
            const long long fw_t0 = clock64();
#endif
            for (int b = 0; b < nb; ++b) {
                const int n = b * FW_MAIN + tid;   // bond index; row index is n + 1
                const bool valid = n < M;
                if (epoch < EPOCH_LOW) {           // rebase the claim epoch (see sweep_cta_kernel)
                    nb_sync(1, FW_MAIN);
                    for (int i = tid; i < (int)claim_n; i += FW_MAIN) claim[i] = CLAIM_FREE;
                    if (tid == 0) sh->star_epoch = 0xffffffffu;
                    epoch = a.epoch_start;
                    nb_sync(1, FW_MAIN);
                }
#ifdef PZ_TIMING
                const long long tw0 = clock64();
#endif
                nb_sync(2, FW_ALL);
#ifdef PZ_TIMING
                const long long tw1 = clock64();
                fw_wait += tw1 - tw0;
#endif
                const uint32_t rep = reps[tid];
                if (b + 1 < nb) nb_arrive(3, FW_ALL);
                uint32_t ru = rep & 0xffffu, rv = rep >> 16, tu = 0, tv = 0;
                if (valid) st.find2(ru, rv, tu, tv);        // up to the current roots
#ifdef PZ_TIMING
                fw_walk += clock64() - tw1;
#endif
                bool pending = valid && ru != rv;
                Rec rec = 0;
                for (bool first = true;; first = false) {
                    // how many bonds of the batch are still pending (the barrier also puts the merges
                    // of the previous round in front of the walks below)
                    const int left = nb_count(1, FW_MAIN, pending);
                    if (left == 0) break;
#ifdef PZ_TIMING
                    const long long fw_r0 = clock64();
#endif
                    if (left <= PZ_TAIL_MAX) {
#ifdef PZ_TIMING
                        ++fw_ntail; fw_titems += left;
#endif
                        // ---- tail mode.  Half of all rounds start with at most 32 pending bonds
                        // (late batches, the losers of a conflict, chains through a large cluster),
                        // and a CTA round costs three barriers over 16 warps whatever the number of
                        // bonds.  The pending bonds are packed, in batch order, into warp 0, which
                        // finishes the batch with warp-level rounds (same rules: claims, owners merge,
                        // star bonds below the first blocked bond merge together) while the others wait.
                        const uint32_t bal = __ballot_sync(0xffffffffu, pending);
                        if (lane == 0) ts->wcnt[warp] = __popc(bal);
                        nb_sync(1, FW_MAIN);
                        int rank = -1;
                        if (pending) {
                            int base = 0;
                            for (int w = 0; w < warp; ++w) base += (int)ts->wcnt[w];
                            rank = base + __popc(bal & ((1u << lane) - 1u));
                            ts->iu[rank] = ru; ts->iv[rank] = rv; ts->in[rank] = (uint32_t)n;
                        }
                        nb_sync(1, FW_MAIN);
                        if (warp == 0) {
                            const bool act = lane < left;
                            uint32_t xu = act ? ts->iu[lane] : 0u, xv = act ? ts->iv[lane] : 0u;
                            const uint32_t xn = act ? ts->in[lane] : 0u;
                            uint32_t yu = 0, yv = 0;
                            Rec xrec = 0;
                            bool xp = act;
                            for (;;) {
                                if (xp) { st.find2(xu, xv, yu, yv); xp = xu != xv; }
                                if (!__any_sync(0xffffffffu, xp)) break;
                                const uint32_t key = (epoch << 10) | (uint32_t)lane;
                                const uint32_t hub = (uint32_t)sh->hub_key;
                                const bool star = xp && (xu == hub || xv == hub);
                                const uint32_t o = xu == hub ? xv : xu, to = xu == hub ? yv : yu;
                                const uint32_t th = xu == hub ? yu : yv;
                                uint32_t su = 0, sv = 0;
                                if (xp) {
                                    su = claim_slot(star ? o : xu, clog);
                                    sv = claim_slot(star ? o : xv, clog);
                                    if (su >= claim_n) su -= claim_n;
                                    if (sv >= claim_n) sv -= claim_n;
                                    atomicMin(&claim[su], key);
                                    if (!star) atomicMin(&claim[sv], key);
                                }
                                __syncwarp();
                                const bool own = xp && claim[su] == key && claim[sv] == key;
                                const uint32_t blocked = __ballot_sync(0xffffffffu, xp && !own);
                                const int lb = blocked ? __ffs(blocked) - 1 : 32;
                                if (own && !star) {
                                    xrec = make_rec<Rec>(Store::size_m1(yu), Store::size_m1(yv));
                                    const uint32_t sz = Store::size_m1(yu) + Store::size_m1(yv) + 2;
                                    const uint32_t m = st.unite(xu, yu, xv, yv, track);
                                    if (track && (m == 3u || a.any3)) atomicMin(&sh->span_min, xn + 1);
                                    const uint32_t big = Store::size_m1(yu) >= Store::size_m1(yv) ? xu : xv;
                                    const unsigned long long hk = ((unsigned long long)sz << 32) | big;
                                    if (hk > sh->hub_key) atomicMax(&sh->hub_key, hk);
                                    xp = false;
                                }
                                const bool sw = own && star && lane < lb;
                                if (__any_sync(0xffffffffu, sw)) {
                                    const uint32_t so = sw ? st.sides_of_root(o, to) : 0u;
                                    const unsigned long long v =
                                        sw ? ((unsigned long long)(Store::size_m1(to) + 1) |
                                              ((unsigned long long)(so & 1u) << 40) |
                                              ((unsigned long long)(so >> 1) << 50)) : 0ull;
                                    unsigned long long incl = v;
#pragma unroll
                                    for (int k = 1; k < 32; k <<= 1) {
                                        const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, k);
                                        if (lane >= k) incl += t;
                                    }
                                    const unsigned long long total = __shfl_sync(0xffffffffu, incl, 31);
                                    if (sw) {
                                        const unsigned long long pre = incl - v;
                                        const uint32_t hub_m1 = Store::size_m1(th);
                                        const uint32_t pre_sz = (uint32_t)(pre & 0xffffffffffull);
                                        xrec = make_rec<Rec>(Store::size_m1(to), hub_m1 + pre_sz);
                                        st.make_child(o, hub);
                                        if (track) {
                                            const uint32_t hs = st.sides_of_root(hub, th);
                                            const uint32_t m = hs | (((incl >> 40) & 0x3ffu) ? 1u : 0u) |
                                                               (((incl >> 50) & 0x3ffu) ? 2u : 0u);
                                            if (m == 3u || a.any3) atomicMin(&sh->span_min, xn + 1);
                                        }
                                        if (pre_sz == 0) {
                                            const uint32_t tot_sz = (uint32_t)(total & 0xffffffffffull);
                                            const uint32_t add = (((total >> 40) & 0x3ffu) ? 1u : 0u) |
                                                                 (((total >> 50) & 0x3ffu) ? 2u : 0u);
                                            st.set_root(hub, hub_m1 + tot_sz, track ? add : 0u);
                                            // the star phase has one writer of the hub: plain compare and store, no CAS loop
                                            const unsigned long long nk = ((unsigned long long)(hub_m1 + tot_sz + 1) << 32) | hub;
                                            if (nk > sh->hub_key) sh->hub_key = nk;
                                        }
                                        xp = false;
                                    }
                                }
                                --epoch;
                                __syncwarp();
#ifdef PZ_TIMING
                                ++fw_trounds;
#endif
                            }
                            if (act) ts->irec[lane] = (unsigned long long)xrec;
                            if (lane == 0) ts->epoch = epoch;
                        }
                        nb_sync(1, FW_MAIN);
                        epoch = ts->epoch;
                        if (rank >= 0) rec = (Rec)ts->irec[rank];
                        pending = false;
#ifdef PZ_TIMING
                        fw_tail += clock64() - fw_r0;
#endif
                        break;
                    }
                    if (!first && pending) {                        // walk up to the new roots
                        st.find2(ru, rv, tu, tv);
                        pending = ru != rv;
                    }
                    // ---- CTA round.  A warp without pending bonds only takes part in the barriers
                    const bool warp_has = __any_sync(0xffffffffu, pending);
                    const uint32_t key = (epoch << 10) | (uint32_t)tid;
                    uint32_t hub = 0, o = 0, to = 0, th = 0, su = 0, sv = 0;
                    bool star = false;
                    if (warp_has) {
                        hub = (uint32_t)sh->hub_key;
                        // star bond: one side is the hub; o = the other root
                        star = pending && (ru == hub || rv == hub);
                        o = ru == hub ? rv : ru; to = ru == hub ? tv : tu;
                        th = ru == hub ? tu : tv;
                        if (pending) {
                            su = claim_slot(star ? o : ru, clog);
                            sv = claim_slot(star ? o : rv, clog);
                            if (su >= claim_n) su -= claim_n;
                            if (sv >= claim_n) sv -= claim_n;
                            atomicMin(&claim[su], key);
                            if (!star) atomicMin(&claim[sv], key);
                            else sh->star_epoch = epoch;            // this round needs the star barrier
                        }
                    }
                    nb_sync(1, FW_MAIN);                            // claims posted
                    const bool star_round = sh->star_epoch == epoch;
                    bool own = false;
                    if (warp_has) {
                        own = pending && claim[su] == key && claim[sv] == key;
                        if (star_round && pending && !own) atomicMin(&sh->bmin, (uint32_t)tid);
                    }
                    bool won = false;
                    if (own && !star) {
                        rec = make_rec<Rec>(Store::size_m1(tu), Store::size_m1(tv));
                        const uint32_t sz = Store::size_m1(tu) + Store::size_m1(tv) + 2;
                        const uint32_t m = st.unite(ru, tu, rv, tv, track);
                        if (track && (m == 3u || a.any3)) atomicMin(&sh->span_min, (uint32_t)n + 1);
                        const uint32_t big = Store::size_m1(tu) >= Store::size_m1(tv) ? ru : rv;
                        const unsigned long long hk = ((unsigned long long)sz << 32) | big;
                        if (hk > sh->hub_key) atomicMax(&sh->hub_key, hk);
                        won = true;
                    }
                    int nstar = 0;
                    if (star_round) nstar = nb_count(1, FW_MAIN, own && star);
                    if (nstar) {
                        // star bonds merge together iff no earlier bond of the batch is blocked
                        const bool sw = own && star && (uint32_t)tid < sh->bmin;
                        unsigned long long v = 0ull, incl = 0ull;
                        if (__any_sync(0xffffffffu, sw)) {
                            const uint32_t so = sw ? st.sides_of_root(o, to) : 0u;
                            v = sw ? ((unsigned long long)(Store::size_m1(to) + 1) |
                                      ((unsigned long long)(so & 1u) << 40) |
                                      ((unsigned long long)(so >> 1) << 50)) : 0ull;
                            incl = v;
#pragma unroll
                            for (int k = 1; k < 32; k <<= 1) {
                                const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, k);
                                if (lane >= k) incl += t;
                            }
                        }
                        if (lane == 31) sh->scan_tot[warp] = incl;
                        nb_sync(1, FW_MAIN);
                        if (sw) {
                            unsigned long long pre = incl - v, total = 0;
#pragma unroll
                            for (int w = 0; w < CTA_WARPS; ++w) {
                                const unsigned long long t = sh->scan_tot[w];
                                if (w < warp) pre += t;
                                total += t;
                            }
                            const uint32_t hub_m1 = Store::size_m1(th);     // hub size - 1 at round start
                            const uint32_t pre_sz = (uint32_t)(pre & 0xffffffffffull);
                            rec = make_rec<Rec>(Store::size_m1(to), hub_m1 + pre_sz);
                            st.make_child(o, hub);
                            if (track) {
                                const uint32_t hs = st.sides_of_root(hub, th);
                                const unsigned long long in = pre + v;
                                const uint32_t m = hs | (((in >> 40) & 0x3ffu) ? 1u : 0u) |
                                                   (((in >> 50) & 0x3ffu) ? 2u : 0u);
                                if (m == 3u || a.any3) atomicMin(&sh->span_min, (uint32_t)n + 1);
                            }
                            if (pre_sz == 0) {      // first star bond of the round: publish the hub
                                const uint32_t tot_sz = (uint32_t)(total & 0xffffffffffull);
                                const uint32_t add = (((total >> 40) & 0x3ffu) ? 1u : 0u) |
                                                     (((total >> 50) & 0x3ffu) ? 2u : 0u);
                                st.set_root(hub, hub_m1 + tot_sz, track ? add : 0u);
                                // the star phase has one writer of the hub: plain compare and store, no CAS loop
                                const unsigned long long nk = ((unsigned long long)(hub_m1 + tot_sz + 1) << 32) | hub;
                                if (nk > sh->hub_key) sh->hub_key = nk;
                            }
                            won = true;
                        }
                    }
                    if (won) pending = false;
                    --epoch;
                    if (tid == 0) sh->bmin = 0xffffffffu;
#ifdef PZ_TIMING
                    fw_cta += clock64() - fw_r0; ++fw_ncta;
#endif
                }
                if (track && sh->span_min != NSPAN_NEVER) track = false;
                if (valid) __stcs(&rec_out[n], rec);
            }
#ifdef PZ_TIMING
            if ((tid & 127) == 0 && blockIdx.x == 0 && run < (int)gridDim.x)
                printf("fw warp %2d: total %lld, waiting for representatives %lld, walking them up %lld; %lld CTA rounds %lld cycles; %lld tails (%lld items, %lld warp rounds) %lld cycles\n",
                       warp, clock64() - fw_t0, fw_wait, fw_walk, fw_ncta, fw_cta, fw_ntail, fw_titems, fw_trounds, fw_tail);
#endif
        }
        __syncthreads();
        if (tid == 0) a.nspan[run] = sh->span_min;
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------
// CTA kernel with finder warps, CARRIED bonds and several HUBS (the shape used for one run per SM).
//
// sweep_fw_kernel finishes every batch before it starts the next: once at most 32 bonds are left,
// warp 0 alone runs warp-level rounds while 15 warps wait (a fifth of all warp samples), and in
// the critical window (0.44 < n/M < 0.51), where several large clusters compete, star merging
// around the single largest cluster leaves chains that take one round per bond.  Here
//
//  * the (at most 32) bonds still pending when a batch has quietened down are CARRIED into the
//    next batch as its earliest bonds: a seventeenth main warp, warp 0, holds nothing but carried
//    bonds (claim keys are the thread index, so they order before every new bond), the other 16
//    main warps take the 512 new bonds of the batch.  There is no tail mode, no warp waits for
//    another, and the rounds of the carried bonds are shared with the next batch's;
//  * star merging works around up to K hubs at once (K = 2 by default): a star bond joins a
//    non-hub root it owns to hub h, all star bonds of a round below the first blocked bond merge
//    together, and ONE ordered 64-bit prefix sum (a 21-bit field per hub) gives each the hub
//    size the sequential order would see.  A bond between two hubs is an ordinary bond that
//    claims both hub roots; it must not overtake an earlier pending star bond of either hub, and
//    a star bond must not overtake an earlier hub-hub bond of its hub (it checks the hub's claim
//    slot without claiming it).  Any set of current roots is a valid hub set; the slots hold the
//    largest clusters seen (a merged cluster replaces the smallest slot it beats, secondary
//    hubs from 256 nodes on).
//
// Simulated with exact claims (scripts/sim_rounds_hubs_carry.c, records identical to the
// sequential run): 427 CTA rounds + 213 tails per L = 256 run -> about 465 CTA rounds, no tails.
// Everything else (finder warps, release/acquire hand-over of the forest, count-first rounds,
// epoch-keyed claims) is sweep_fw_kernel's.
// ---------------------------------------------------------------------------
static constexpr int FC_NEW_WARPS = 16, FC_MAIN_WARPS = FC_NEW_WARPS + 1;
static constexpr int FC_NEW = 32 * FC_NEW_WARPS, FC_MAIN = 32 * FC_MAIN_WARPS;       // 512, 544
static constexpr int FC_FIND = FW_FIND, FC_ALL = FC_MAIN + FC_FIND;                   // 128, 672
static constexpr int FC_PAIR = FC_NEW + FC_FIND;                                      // hand-over barriers
static constexpr int FC_BPT = FC_NEW / FC_FIND;
static constexpr int FC_MAX_HUBS = 3;
#ifndef PZ_HUB_MIN
#define PZ_HUB_MIN 256                        // smallest cluster that may become a secondary hub
#endif

struct FcShared {
    unsigned long long hubk[FC_MAX_HUBS];     // (size << 32) | root of hub k, 0 = empty slot
    unsigned long long scan_tot[FC_MAIN_WARPS];
    uint32_t scan_or[FC_MAIN_WARPS];
    uint32_t firststar[FC_MAX_HUBS];          // epoch-keyed: lowest pending star bond of hub k
    uint32_t span_min;
    uint32_t bmin;                            // epoch-keyed: lowest blocked bond of the round
    uint32_t star_epoch;
    uint32_t wcnt[FC_MAIN_WARPS];             // pending bonds per main warp
    uint32_t iu[32], iv[32], in[32];          // carried bonds, in order
};
static_assert(sizeof(FcShared) <= 320 + 768, "FcShared must fit the planner's reserve");

#ifdef PZ_TIMING
// clock read that waits for `dep` (a value produced by the operation being timed)
__device__ __forceinline__ long long tm_now(uint32_t dep) {
    long long t;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) : "r"(dep) : "memory");
    return t;
}
#define FCT(...) __VA_ARGS__
#else
#define FCT(...)
#endif

template <class Store, int K>
__global__ void __launch_bounds__(FC_ALL + 256, 1) sweep_fc_kernel(SweepArgs a, uint32_t store_bytes)
{
    extern __shared__ __align__(16) unsigned char smem[];
    using Rec = typename Store::Rec;
    using Edge = typename Store::Edge;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool finder = tid >= FC_MAIN;
    const bool carrier = warp == 0;
    const int M = a.M;
    const int clog = a.claim_log2;

    uint32_t *claim = reinterpret_cast<uint32_t *>(smem);
    const uint32_t claim_n = (1u << clog) - FC_NEW;
    uint32_t *reps = claim + claim_n;                                          // [FC_NEW] ru | rv << 16
    unsigned char *p = smem + (sizeof(uint32_t) << clog);
    Store st;
    st.bind(p, a, blockIdx.x);
    FcShared *sh = reinterpret_cast<FcShared *>(p + store_bytes);

    const Edge *edges = reinterpret_cast<const Edge *>(a.edges);
    const bool spanning = a.sides2 != nullptr;
    const int nb = (M + FC_NEW - 1) / FC_NEW;

    for (int run = blockIdx.x; run < a.R; run += gridDim.x) {
        st.init(tid, FC_ALL);
        for (int i = tid; i < (int)claim_n; i += FC_ALL) claim[i] = CLAIM_FREE;
        if (tid == 0) {
            sh->span_min = NSPAN_NEVER;
            sh->bmin = 0xffffffffu;
            sh->star_epoch = 0xffffffffu;
#pragma unroll
            for (int k = 0; k < FC_MAX_HUBS; ++k) { sh->hubk[k] = 0ull; sh->firststar[k] = 0xffffffffu; }
        }
        __syncthreads();
        const int32_t *perm = a.perms + (size_t)run * M;

        if (finder) {
            // ---- finder warps: FC_BPT bonds per thread and batch (as in sweep_fw_kernel) -------
            const int ft = tid - FC_MAIN;
            Edge uv_next[FC_BPT];
            int32_t e_next[FC_BPT];
#pragma unroll
            for (int q = 0; q < FC_BPT; ++q) {
                uv_next[q] = Edge(); e_next[q] = 0;
                const int n = ft + q * FC_FIND;
                if (n < M) uv_next[q] = __ldg(&edges[__ldcs(&perm[n])]);
                if (n + FC_NEW < M) e_next[q] = __ldcs(&perm[n + FC_NEW]);
            }
            for (int b = 0; b < nb; ++b) {
                uint32_t out[FC_BPT];
#pragma unroll
                for (int q = 0; q < FC_BPT; ++q) {
                    const int n = b * FC_NEW + ft + q * FC_FIND;
                    const Edge uv = uv_next[q];
                    if (n + FC_NEW < M) uv_next[q] = __ldg(&edges[e_next[q]]);
                    if (n + 2 * FC_NEW < M) e_next[q] = __ldcs(&perm[n + 2 * FC_NEW]);
                    uint32_t u = 0, v = 0;
                    if (n < M) {
                        edge_uv(uv, u, v);
                        uint32_t x[2] = {u, v};
                        st.template find_rep_multi<2>(x);
                        u = x[0]; v = x[1];
                    }
                    out[q] = u | (v << 16);
                }
                nb_sync(3, FC_PAIR);                    // the buffer of the previous batch has been read
#pragma unroll
                for (int q = 0; q < FC_BPT; ++q) reps[ft + q * FC_FIND] = out[q];
                nb_arrive(2, FC_PAIR);                  // representatives of batch b are in the buffer
            }
        } else {
            // ---- main warps ---------------------------------------------------------------------
            Rec *rec_out = reinterpret_cast<Rec *>(a.recs) + (size_t)run * M;
            bool track = spanning;                 // uniform over the main warps
            uint32_t epoch = a.epoch_start;
            if (nb > 0 && !carrier) nb_arrive(3, FC_PAIR);      // the buffer is free
            // the bond this thread holds: a new thread one bond per batch, a carrier lane a carried one
            uint32_t ru = 0, rv = 0, tu = 0, tv = 0;
            int n = 0;
            bool has = false, pending = false;
            Rec rec = 0;
            FCT(long long t_reps = 0, t_walk = 0, t_count = 0, t_round = 0, t_pack = 0, n_round = 0, n_star = 0, n_pack = 0, n_left = 0, t_star = 0;
                const long long t_begin = tm_now(0);)
            for (int b = 0; b < nb; ++b) {
                const bool last = b + 1 == nb;
                if (epoch < EPOCH_LOW) {           // rebase the claim epoch (see sweep_cta_kernel)
                    nb_sync(1, FC_MAIN);
                    for (int i = tid; i < (int)claim_n; i += FC_MAIN) claim[i] = CLAIM_FREE;
                    if (tid == 0) {
                        sh->star_epoch = 0xffffffffu; sh->bmin = 0xffffffffu;
#pragma unroll
                        for (int k = 0; k < FC_MAX_HUBS; ++k) sh->firststar[k] = 0xffffffffu;
                    }
                    epoch = a.epoch_start;
                    nb_sync(1, FC_MAIN);
                }
                FCT(const long long tq0 = tm_now(0);)
                if (!carrier) {
                    n = b * FC_NEW + (tid - 32);
                    has = n < M;
                    nb_sync(2, FC_PAIR);
                    const uint32_t rep = reps[tid - 32];
                    FCT(t_reps += tm_now(rep) - tq0;)
                    if (!last) nb_arrive(3, FC_PAIR);
                    ru = rep & 0xffffu; rv = rep >> 16;
                    rec = 0;
                    pending = has;
                }
                FCT(const long long tq1 = tm_now(0);)
                if (pending) {                      // up to the current roots
                    st.find2(ru, rv, tu, tv);
                    pending = ru != rv;
                }
                FCT(t_walk += tm_now(ru ^ rv) - tq1;)
                bool carried_away = false;          // my bond moved to the carrier warp: it writes the record
                int left = 0;
                for (bool first = true;; first = false) {
                    const uint32_t bal = __ballot_sync(0xffffffffu, pending);
                    if (lane == 0) sh->wcnt[warp] = __popc(bal);
                    // how many bonds are pending (the barrier also puts the merges of the previous
                    // round in front of the walks below)
                    FCT(const long long tq2 = tm_now(0);)
                    left = nb_count(1, FC_MAIN, pending);
                    FCT(const long long tq3 = tm_now((uint32_t)left); t_count += tq3 - tq2;)
                    if (left == 0) break;
                    if (left <= 32 && !last) {
                        // ---- carry: pack the pending bonds, in order, into the carrier warp -------
                        if (carrier && has && !pending) { __stcs(&rec_out[n], rec); has = false; }
                        if (pending) {
                            int base = 0;
                            for (int w = 0; w < warp; ++w) base += (int)sh->wcnt[w];
                            const int rank = base + __popc(bal & ((1u << lane) - 1u));
                            sh->iu[rank] = ru; sh->iv[rank] = rv; sh->in[rank] = (uint32_t)n;
                            carried_away = !carrier;
                        }
                        nb_sync(1, FC_MAIN);
                        if (carrier) {
                            has = pending = lane < left;
                            rec = 0;
                            if (has) { ru = sh->iu[lane]; rv = sh->iv[lane]; n = (int)sh->in[lane]; }
                        } else {
                            pending = false;
                        }
                        FCT(t_pack += tm_now(ru) - tq3; ++n_pack; n_left += left;)
                        break;
                    }
                    if (!first && pending) {                        // walk up to the new roots
                        st.find2(ru, rv, tu, tv);
                        pending = ru != rv;
                    }
                    // ---- CTA round.  A warp without pending bonds only takes part in the barriers
                    const bool warp_has = __any_sync(0xffffffffu, pending);
                    const uint32_t key = (epoch << 10) | (uint32_t)tid;
                    uint32_t hubroot = 0, o = 0, to = 0, th = 0, su = 0, sv = 0;
                    int hu = -1, hv = -1, h = 0;
                    unsigned long long hmin = ~0ull, hub0 = 0ull;   // smallest hub slot (candidates replace it)
                    int kmin = 0;
                    bool star = false;
                    if (warp_has) {
#pragma unroll
                        for (int k = 0; k < K; ++k) {
                            const unsigned long long hk = sh->hubk[k];
                            if (k == 0) hub0 = hk;
                            if (hk < hmin) { hmin = hk; kmin = k; }
                            if (hk != 0ull) {
                                if (ru == (uint32_t)hk) hu = k;
                                if (rv == (uint32_t)hk) hv = k;
                            }
                        }
                        // star bond: exactly one side is a hub; o = the other root
                        star = pending && ((hu >= 0) != (hv >= 0));
                        h = hu >= 0 ? hu : hv;
                        hubroot = hu >= 0 ? ru : rv;
                        o = hu >= 0 ? rv : ru; to = hu >= 0 ? tv : tu;
                        th = hu >= 0 ? tu : tv;
                        if (pending) {
                            su = claim_slot(star ? o : ru, clog);
                            sv = claim_slot(star ? hubroot : rv, clog);
                            if (su >= claim_n) su -= claim_n;
                            if (sv >= claim_n) sv -= claim_n;
                            atomicMin(&claim[su], key);
                            if (!star) atomicMin(&claim[sv], key);
                            else {
                                if (K > 1) atomicMin(&sh->firststar[h], key);
                                sh->star_epoch = epoch;             // this round needs the star barrier
                            }
                        }
                    }
                    nb_sync(1, FC_MAIN);                            // claims posted
                    const bool star_round = sh->star_epoch == epoch;
                    bool own = false;
                    if (warp_has && pending) {
                        const uint32_t cu = claim[su], cv = claim[sv];
                        // a star bond owns its other root and no earlier bond claims its hub
                        own = cu == key && (star ? (K == 1 || cv >= key) : cv == key);
                        // a bond between two hubs waits for the earlier star bonds of both
                        if (K > 1 && own && !star && hu >= 0)
                            own = sh->firststar[hu] > key && sh->firststar[hv] > key;
                        if (star_round && !own) atomicMin(&sh->bmin, key);
                    }
                    bool won = false;
                    if (own && !star) {
                        const uint32_t a1 = Store::size_m1(tu), b1 = Store::size_m1(tv);
                        rec = make_rec<Rec>(a1, b1);
                        const uint32_t sz = a1 + b1 + 2;
                        const uint32_t m = st.unite(ru, tu, rv, tv, track);
                        if (track && (m == 3u || a.any3)) atomicMin(&sh->span_min, (uint32_t)n + 1);
                        const uint32_t big = a1 >= b1 ? ru : rv;
                        const unsigned long long hk = ((unsigned long long)sz << 32) | big;
                        if (hu >= 0) {                              // two hubs became one
                            const int hb = a1 >= b1 ? hu : hv, hs = a1 >= b1 ? hv : hu;
                            atomicMax(&sh->hubk[hb], hk);
                            sh->hubk[hs] = 0ull;
                        } else if (hk > hmin && (kmin == 0 || sz >= PZ_HUB_MIN)) {
                            atomicMax(&sh->hubk[kmin], hk);
                        } else if (K > 1 && hk > hub0) {            // too small for a secondary hub: slot 0 takes any size
                            atomicMax(&sh->hubk[0], hk);
                        }
                        won = true;
                    }
                    int nstar = 0;
                    if (star_round) nstar = nb_count(1, FC_MAIN, own && star);
                    if (nstar) {
                        // star bonds merge together iff no earlier bond is blocked
                        const bool sw = own && star && key < sh->bmin;
                        unsigned long long v = 0ull, incl = 0ull;
                        uint32_t sb = 0u, sincl = 0u, sexcl = 0u;
                        if (__any_sync(0xffffffffu, sw)) {
                            if (sw) {
                                v = (unsigned long long)(Store::size_m1(to) + 1) << (21 * h);
                                if (track) sb = st.sides_of_root(o, to) << (2 * h);
                            }
                            incl = v; sincl = sb;
#pragma unroll
                            for (int k = 1; k < 32; k <<= 1) {
                                const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, k);
                                const uint32_t t2 = __shfl_up_sync(0xffffffffu, sincl, k);
                                if (lane >= k) { incl += t; sincl |= t2; }
                            }
                            sexcl = __shfl_up_sync(0xffffffffu, sincl, 1);
                            if (lane == 0) sexcl = 0u;
                        }
                        if (lane == 31) { sh->scan_tot[warp] = incl; sh->scan_or[warp] = sincl; }
                        nb_sync(1, FC_MAIN);
                        if (sw) {
                            unsigned long long pre = incl - v, total = 0ull;
                            uint32_t spre = sexcl, stot = 0u;
#pragma unroll
                            for (int w = 0; w < FC_MAIN_WARPS; ++w) {
                                const unsigned long long t = sh->scan_tot[w];
                                const uint32_t t2 = sh->scan_or[w];
                                if (w < warp) { pre += t; spre |= t2; }
                                total += t; stot |= t2;
                            }
                            const uint32_t hub_m1 = Store::size_m1(th);     // hub size - 1 at round start
                            const uint32_t pre_sz = (uint32_t)(pre >> (21 * h)) & 0x1fffffu;
                            rec = make_rec<Rec>(Store::size_m1(to), hub_m1 + pre_sz);
                            st.make_child(o, hubroot);
                            if (track) {
                                const uint32_t m = st.sides_of_root(hubroot, th) |
                                                   (((spre | sb) >> (2 * h)) & 3u);
                                if (m == 3u || a.any3) atomicMin(&sh->span_min, (uint32_t)n + 1);
                            }
                            if (pre_sz == 0) {      // first star bond of its hub in this round: publish the hub
                                const uint32_t tot_sz = (uint32_t)(total >> (21 * h)) & 0x1fffffu;
                                st.set_root(hubroot, hub_m1 + tot_sz, track ? ((stot >> (2 * h)) & 3u) : 0u);
                                const unsigned long long nk = ((unsigned long long)(hub_m1 + tot_sz + 1) << 32) | hubroot;
                                if (nk > sh->hubk[h]) sh->hubk[h] = nk;
                            }
                            won = true;
                        }
                    }
                    if (won) pending = false;
                    --epoch;
                    FCT(t_round += tm_now((uint32_t)rec) - tq3; ++n_round; if (nstar) ++n_star;)
                }
                // spanning is decided once no earlier bond can still join the two sides
                if (track && left == 0 && sh->span_min != NSPAN_NEVER) track = false;
                if (carrier) {
                    if (has && !pending) { __stcs(&rec_out[n], rec); has = false; }
                } else if (has && !carried_away) {
                    __stcs(&rec_out[n], rec);
                }
            }
            FCT(if ((lane == 0) && (warp == 0 || warp == 1 || warp == 9) && blockIdx.x == 3 && run < 2 * (int)gridDim.x)
                    printf("fc run %d warp %2d: total %lld | reps %lld | walk %lld | count barriers %lld | %lld rounds (%lld star) %lld cycles | %lld packs (%lld bonds) %lld cycles\n",
                           run, warp, tm_now(0) - t_begin, t_reps, t_walk, t_count, n_round, n_star, t_round, n_pack, n_left, t_pack);)
        }
        __syncthreads();
        if (tid == 0) a.nspan[run] = sh->span_min;
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------
// planning and launch
// ---------------------------------------------------------------------------

static int ilog2_ceil(uint32_t x) { int l = 0; while ((1u << l) < x) ++l; return l; }

static SweepPlan plan_team(SweepPlan p, int32_t N, int32_t R, int sms, size_t smem_optin,
                           size_t store_bytes, int claim_cap, int cta_warps)
{
    // one CTA of CTA_THREADS per run; as many CTAs per SM as shared memory and
    // the 64-warp limit allow.  Claim table: exact (one slot per node) when it
    // fits next to the store, else as large as fits (hashed).
    const size_t sm_total = 228 * 1024;
    const size_t fixed = store_bytes + 320;    // + CtaShared
    const int CTA_WARPS = cta_warps;
    int clog = ilog2_ceil((uint32_t)(N < 256 ? 256 : N));
    if (claim_cap < 8) claim_cap = 8;
    if (claim_cap > 14) claim_cap = 14;
    if (clog > claim_cap) clog = claim_cap;
    while (clog > 8 && fixed + ((size_t)4 << clog) > smem_optin) --clog;
    p.claim_log2 = clog;
    p.team = 1;
    p.store_bytes = store_bytes;
    p.warps_per_cta = CTA_WARPS;
    p.smem_bytes = align16h(fixed + ((size_t)4 << clog));
    // one run per SM (shared-memory store, 16 warps): finder warps walk the next batch up the forest
    // while the main warps run the rounds (PZ_FINDERS=0 switches them off)
    p.finders = 0;
    if (p.kind == STORE_S16B && CTA_WARPS == 16) {
        int want = 1;           // 0: lock step, 1: finder warps + tail mode, 2: finder warps + carried bonds + hubs
        if (const char *e = getenv("PZ_FINDERS")) want = atoi(e);
        if (want && clog >= 11 && align16h(fixed + ((size_t)4 << clog) + 768) <= smem_optin) {
            p.finders = want;   // (the hand-over buffer is carved out of the claim table)
            p.smem_bytes = align16h(fixed + ((size_t)4 << clog) + 768);      // + TailShared
        }
    }
    p.slice_bytes = p.smem_bytes;
    int ctas_per_sm = (int)(sm_total / (p.smem_bytes + 1024));
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    if (ctas_per_sm > 64 / CTA_WARPS) ctas_per_sm = 64 / CTA_WARPS;
    {
        // measured (profiles/sweep_cta_shape_r1.txt): 4-warp CTAs of the global-memory store keep
        // gaining up to 6 per SM (L = 256: 1.74e10, 3D L = 64: 1.54e10 bonds/s), 8 is no better
        int cap = 6;                            // PZ_G32_CTAS overrides
        if (const char *e = getenv("PZ_G32_CTAS")) cap = atoi(e);
        if (p.kind == STORE_G32 && ctas_per_sm > cap) ctas_per_sm = cap;
    }
    long long grid = (long long)sms * ctas_per_sm;
    if (grid > R) grid = R;
    if (grid < 1) grid = 1;
    p.grid = (int)grid;
    p.gscratch_bytes = p.kind == STORE_G32 ? (size_t)p.grid * (size_t)N * 4 : 0;
    return p;
}

SweepPlan plan_sweep(int32_t N, int32_t R, int sms, size_t smem_optin, int force_kind, int team,
                     int claim_cap, int cta_warps)
{
    SweepPlan p{};
    const size_t budget = smem_optin;          // per CTA (opt-in maximum)
    StoreKind kind;
    if (force_kind >= 0) kind = (StoreKind)force_kind;
    else if (N <= 32768 && StoreS16::slice_bytes(N) + 1024 <= budget) kind = STORE_S16;
    else if (N <= 65536 && StoreS16B::slice_bytes(N) + 1024 <= budget) kind = STORE_S16B;
    else kind = STORE_G32;
    p.kind = kind;

    size_t store_bytes = kind == STORE_S16 ? StoreS16::slice_bytes(N)
                       : kind == STORE_S16B ? StoreS16B::slice_bytes(N) : 0;
    // claim table: exact (one slot per node) when small, else hashed
    int clog = ilog2_ceil((uint32_t)(N < 64 ? 64 : N));
    const int clog_max = kind == STORE_G32 ? 12 : 13;
    if (clog > clog_max) clog = clog_max;
    while (clog > 8 && store_bytes + ((size_t)4 << clog) > budget) --clog;
    p.claim_log2 = clog;
    if (team) {
        // measured on B200 (profiles/sweep_cta_shape_r1.txt): when only one CTA fits an SM
        // (L = 256) wide batches win (16 warps, largest claim table); when several fit, four
        // warps per CTA and a 16 KB claim table keep more runs in flight
        int w = cta_warps, cap = claim_cap;
        if (w <= 0) {
            const size_t sb = align16h(store_bytes);
            const int fit = (int)((228 * 1024) / (sb + ((size_t)4 << 12) + 1024 + 320));
            w = fit >= 4 ? 4 : fit >= 2 ? 8 : 16;
            if (cap <= 0) cap = fit >= 2 ? 12 : 14;
        }
        if (cap <= 0) cap = 12;
        w = w <= 2 ? 2 : w <= 4 ? 4 : w <= 8 ? 8 : w <= 16 ? 16 : 32;
        return plan_team(p, N, R, sms, smem_optin, align16h(store_bytes), cap, w);
    }
    p.slice_bytes = align16h(store_bytes + ((size_t)4 << clog));

    // warps (runs in flight) per CTA and CTAs per SM: as many runs as shared
    // memory (228 KB per SM, 1 KB reserved per CTA) and 64 warps allow
    const size_t sm_total = 228 * 1024;
    int wpc = 1;
    if (kind == STORE_G32) {
        wpc = 8;
    } else {
        size_t fit = budget / p.slice_bytes;            // warps that fit one CTA
        if (fit < 1) fit = 1;
        wpc = (int)(fit > 8 ? 8 : fit);
    }
    // never more warps than runs need
    long long want = ((long long)R + sms - 1) / sms;
    if (want < 1) want = 1;
    if (wpc > want) wpc = (int)want;
    p.warps_per_cta = wpc;
    p.smem_bytes = p.slice_bytes * (size_t)wpc;
    int ctas_per_sm = (int)(sm_total / (p.smem_bytes + 1024));
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    int by_warps = 64 / wpc;
    if (ctas_per_sm > by_warps) ctas_per_sm = by_warps;
    if (kind == STORE_G32 && ctas_per_sm > 2) ctas_per_sm = 2;
    long long grid = (long long)sms * ctas_per_sm;
    long long need = ((long long)R + wpc - 1) / wpc;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    p.grid = (int)grid;
    p.gscratch_bytes = kind == STORE_G32 ? (size_t)p.grid * wpc * (size_t)N * 4 : 0;
    return p;
}

template <class Store, int W>
static cudaError_t launch_team_w(const SweepPlan &p, const SweepArgs &a, cudaStream_t s)
{
    cudaError_t e = cudaFuncSetAttribute(sweep_cta_kernel<Store, W>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)p.smem_bytes);
    if (e != cudaSuccess) return e;
    sweep_cta_kernel<Store, W><<<p.grid, 32 * W, p.smem_bytes, s>>>(a, (uint32_t)p.store_bytes);
    return cudaGetLastError();
}

template <class Store>
static cudaError_t launch_team_t(const SweepPlan &p, const SweepArgs &a, cudaStream_t s)
{
    switch (p.warps_per_cta) {
    case 2: return launch_team_w<Store, 2>(p, a, s);
    case 4: return launch_team_w<Store, 4>(p, a, s);
    case 8: return launch_team_w<Store, 8>(p, a, s);
    case 32: return launch_team_w<Store, 32>(p, a, s);
    default: return launch_team_w<Store, 16>(p, a, s);
    }
}

template <class Store>
static cudaError_t launch_t(const SweepPlan &p, const SweepArgs &a, cudaStream_t s)
{
    cudaError_t e = cudaFuncSetAttribute(sweep_kernel<Store>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)p.smem_bytes);
    if (e != cudaSuccess) return e;
    sweep_kernel<Store><<<p.grid, p.warps_per_cta * 32, p.smem_bytes, s>>>(a, (uint32_t)p.slice_bytes);
    return cudaGetLastError();
}

cudaError_t launch_sweep(const SweepPlan &p, const SweepArgs &a, cudaStream_t s)
{
    if (p.team && p.finders >= 2) {
        static const int hubs = getenv("PZ_HUBS") ? atoi(getenv("PZ_HUBS")) : 2;
        auto kern = hubs <= 1 ? sweep_fc_kernel<StoreS16BF, 1> : hubs == 2 ? sweep_fc_kernel<StoreS16BF, 2>
                                                                           : sweep_fc_kernel<StoreS16BF, 3>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_bytes);
        if (e != cudaSuccess) return e;
        kern<<<p.grid, FC_ALL, p.smem_bytes, s>>>(a, (uint32_t)p.store_bytes);
        return cudaGetLastError();
    }
    if (p.team && p.finders) {
        cudaError_t e = cudaFuncSetAttribute(sweep_fw_kernel<StoreS16BF>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_bytes);
        if (e != cudaSuccess) return e;
        sweep_fw_kernel<StoreS16BF><<<p.grid, FW_ALL, p.smem_bytes, s>>>(a, (uint32_t)p.store_bytes);
        return cudaGetLastError();
    }
    if (p.team) {
        switch (p.kind) {
        case STORE_S16: return launch_team_t<StoreS16>(p, a, s);
        case STORE_S16B: return launch_team_t<StoreS16B>(p, a, s);
        default: return launch_team_t<StoreG32>(p, a, s);
        }
    }
    switch (p.kind) {
    case STORE_S16: return launch_t<StoreS16>(p, a, s);
    case STORE_S16B: return launch_t<StoreS16B>(p, a, s);
    default: return launch_t<StoreG32>(p, a, s);
    }
}

}  // namespace pz
