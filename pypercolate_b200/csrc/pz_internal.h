// pz_internal.h -- host-side declarations shared by the .cu translation units.
#pragma once
#include <cstdint>
#include <cstddef>
#include <cuda_runtime.h>

namespace pz {

enum StoreKind {
    STORE_S16 = 0,    // shared memory, uint16 parent/size, root flag in bit 15  (N <= 32768)
    STORE_S16B = 1,   // shared memory, uint16 parent/size + root bitmap         (N <= 65536)
    STORE_G32 = 2     // global memory (L2 / HBM resident), uint32               (N <  2^29)
};

struct SweepArgs {
    int32_t N, M, R;
    const void *edges;        // uint32 (u | v << 16) when N <= 65536, else uint2 {u, v}
    const uint32_t *sides2;   // 2-bit side fields, (N + 15) / 16 words; nullptr = no spanning
    int any3;                 // sides already joined by the auxiliary structure alone
    const int32_t *perms;     // [R][M] bond order of every run
    void *recs;               // [R][M] merge records (uint32 when N <= 65536 else uint64)
    uint32_t *nspan;          // [R] first n with a spanning cluster (NSPAN_NEVER if none)
    uint32_t *gscratch;       // STORE_G32: [total warps][N]
    int claim_log2;           // log2 of the per-warp claim table (entries)
    uint32_t epoch_start;     // first claim epoch of a run (22 bits; PZ_EPOCH_START shortens it for tests)
};

struct SweepPlan {
    StoreKind kind;
    int warps_per_cta;
    int grid;
    size_t smem_bytes;        // dynamic shared memory per CTA
    size_t slice_bytes;       // per warp
    int claim_log2;
    size_t gscratch_bytes;
    int team;                 // 1: finder/merger team kernel (one CTA of 4 warps per run)
    size_t store_bytes;       // team kernel: bytes of the parent store inside the CTA's smem
    int finders;              // 1: CTA kernel with 8 finder warps next to 16 main warps
};

// choose store, CTA shape and grid for a graph of N nodes on a device with
// `sms` SMs and `smem_optin` bytes of opt-in shared memory per CTA
SweepPlan plan_sweep(int32_t N, int32_t R, int sms, size_t smem_optin, int force_kind, int team,
                     int claim_cap_log2, int cta_warps);
cudaError_t launch_sweep(const SweepPlan &plan, const SweepArgs &args, cudaStream_t stream);

// ---- statistics kernels (pz_stats.cu) ---------------------------------------
struct StatsArgs {
    int32_t N, M, R;
    int rec64;                // records are uint64
    const void *recs;         // [R][M]
    const uint32_t *nspan;    // [R]
    const int32_t *perms;     // [R][M] (rows only)
    int spanning;
};
cudaError_t launch_expand_rows(const StatsArgs &a, uint8_t *rows, cudaStream_t stream);

}  // namespace pz
