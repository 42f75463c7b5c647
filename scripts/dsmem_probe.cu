// dsmem_probe.cu -- where could the parent array of a large run live?  (BASELINE config 4: L = 1024, N = 2^20 nodes)
// Dependent-load chains (the access pattern of a union-find walk: the next index is the value just loaded) over a
// table of 3 MiB (2^20 nodes at 24 bits) in
//   (a) the distributed shared memory of a 16-CTA cluster (192 KB per CTA, ld.shared::cluster through mapa),
//   (b) global memory, small enough to stay in the L2,
//   (c) global memory among 3.5 GB of such tables (888 runs in flight: what ships), i.e. HBM,
// at 1 .. 16 warps per CTA, and the cost of the barrier a lock-step round needs (cluster vs CTA).
// Build + run (on the GPU box):  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/dsmem_probe scripts/dsmem_probe.cu && /tmp/dsmem_probe
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

static constexpr int CL = 16;                     // CTAs per cluster
static constexpr int PER = 48 * 1024;             // uint32 entries per CTA (192 KB)
static constexpr int TOTAL = CL * PER;            // 786432 entries = 3 MiB

__device__ __forceinline__ uint32_t ld_cluster(uint32_t local_base, uint32_t idx)
{
    // entry idx lives in CTA idx / PER at offset idx % PER
    const uint32_t cta = idx / PER, off = idx % PER;
    uint32_t remote, v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_base + off * 4u), "r"(cta));
    asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(remote) : "memory");
    return v;
}

__global__ void dsmem_chase(const uint32_t *table, int steps, long long *cycles, uint32_t *sink)
{
    extern __shared__ __align__(16) uint32_t sm[];
    cg::cluster_group cluster = cg::this_cluster();
    const uint32_t rank = cluster.block_rank();
    for (int i = threadIdx.x; i < PER; i += blockDim.x) sm[i] = table[rank * PER + i];
    cluster.sync();
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm);
    uint32_t idx = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u % TOTAL;
    const long long t0 = clock64();
    for (int s = 0; s < steps; ++s) idx = ld_cluster(base, idx);
    const long long t1 = clock64();
    if (threadIdx.x == 0 && rank == 0) cycles[blockIdx.x / CL] = t1 - t0;
    if (idx == 0xffffffffu) *sink = idx;
    cluster.sync();
}

__global__ void local_chase(const uint32_t *table, int steps, long long *cycles, uint32_t *sink)
{
    extern __shared__ __align__(16) uint32_t sm[];
    for (int i = threadIdx.x; i < PER; i += blockDim.x) sm[i] = table[i] % PER;
    __syncthreads();
    uint32_t idx = (threadIdx.x * 2654435761u) % PER;
    const long long t0 = clock64();
    for (int s = 0; s < steps; ++s) idx = sm[idx];
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (idx == 0xffffffffu) *sink = idx;
}

__global__ void global_chase(const uint32_t *tables, size_t stride, int steps, long long *cycles, uint32_t *sink)
{
    const uint32_t *t = tables + (size_t)blockIdx.x * stride;
    uint32_t idx = (threadIdx.x * 2654435761u + blockIdx.x * 40503u) % TOTAL;
    const long long t0 = clock64();
    for (int s = 0; s < steps; ++s) idx = t[idx];
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (idx == 0xffffffffu) *sink = idx;
}

__global__ void cluster_barrier_cost(int n, long long *cycles)
{
    cg::cluster_group cluster = cg::this_cluster();
    cluster.sync();
    const long long t0 = clock64();
    for (int i = 0; i < n; ++i) cluster.sync();
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}
__global__ void cta_barrier_cost(int n, long long *cycles)
{
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < n; ++i) __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

static double avg(const std::vector<long long> &v) { double s = 0; for (auto x : v) s += (double)x; return s / v.size(); }

int main()
{
    // one random cycle through all entries: every load depends on the one before
    std::vector<uint32_t> perm(TOTAL), table(TOTAL);
    for (int i = 0; i < TOTAL; ++i) perm[i] = i;
    uint64_t r = 88172645463325252ull;
    for (int i = TOTAL - 1; i > 0; --i) { r ^= r << 13; r ^= r >> 7; r ^= r << 17; std::swap(perm[i], perm[r % (i + 1)]); }
    for (int i = 0; i < TOTAL; ++i) table[perm[i]] = perm[(i + 1) % TOTAL];
    const int steps = 2000;
    uint32_t *d_table, *d_sink; long long *d_cyc;
    CK(cudaMalloc(&d_table, TOTAL * 4)); CK(cudaMalloc(&d_sink, 4)); CK(cudaMalloc(&d_cyc, 8 * 4096));
    CK(cudaMemcpy(d_table, table.data(), TOTAL * 4, cudaMemcpyHostToDevice));
    std::vector<long long> h(4096);

    printf("dependent-load chains over a 3 MiB table (%d steps per thread), cycles per step seen by one thread\n", steps);
    printf("%-44s %8s %8s %8s %8s %8s\n", "warps per CTA", "1", "2", "4", "8", "16");
    {   // (0) local shared memory, one CTA's own 192 KB
        CK(cudaFuncSetAttribute(local_chase, cudaFuncAttributeMaxDynamicSharedMemorySize, PER * 4));
        printf("%-44s", "shared memory of the CTA itself (192 KB)");
        for (int w : {1, 2, 4, 8, 16}) {
            local_chase<<<148, 32 * w, PER * 4>>>(d_table, steps, d_cyc, d_sink);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(h.data(), d_cyc, 8 * 148, cudaMemcpyDeviceToHost));
            h.resize(148); printf(" %8.1f", avg(h) / steps); h.resize(4096);
        }
        printf("\n");
    }
    {   // (a) DSMEM of a 16-CTA cluster, 9 clusters
        CK(cudaFuncSetAttribute(dsmem_chase, cudaFuncAttributeMaxDynamicSharedMemorySize, PER * 4));
        CK(cudaFuncSetAttribute(dsmem_chase, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        printf("%-44s", "distributed shared memory, 16-CTA cluster");
        for (int w : {1, 2, 4, 8, 16}) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(9 * CL); cfg.blockDim = dim3(32 * w); cfg.dynamicSmemBytes = PER * 4;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            cudaError_t e = cudaLaunchKernelEx(&cfg, dsmem_chase, (const uint32_t *)d_table, steps, d_cyc, d_sink);
            if (e != cudaSuccess) { printf(" %8s", cudaGetErrorString(e)); cudaGetLastError(); continue; }
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(h.data(), d_cyc, 8 * 9, cudaMemcpyDeviceToHost));
            h.resize(9); printf(" %8.1f", avg(h) / steps); h.resize(4096);
        }
        printf("   (9 clusters = 144 SMs)\n");
    }
    for (int pass = 0; pass < 3; ++pass) {
        // (b) 24 tables (72 MB: L2 resident), (c) 148, (d) 888 tables of 4 MiB stride (3.5 GB: HBM)
        const int ntab = pass == 0 ? 24 : pass == 1 ? 148 : 888;
        const size_t stride = (size_t)1 << 20;        // 4 MiB apart, like the uint32 parent arrays of L = 1024
        uint32_t *d_tabs;
        CK(cudaMalloc(&d_tabs, (size_t)ntab * stride * 4));
        for (int k = 0; k < ntab; ++k) CK(cudaMemcpy(d_tabs + k * stride, d_table, TOTAL * 4, cudaMemcpyDeviceToDevice));
        char name[96];
        snprintf(name, sizeof name, "global memory, %d runs in flight (%.1f GB)", ntab, ntab * stride * 4 / 1e9);
        printf("%-44s", name);
        for (int w : {1, 2, 4, 8, 16}) {
            if (pass == 2 && w > 4) { printf(" %8s", "-"); continue; }      // (888 CTAs of 4 warps is the shipped shape)
            global_chase<<<ntab, 32 * w>>>(d_tabs, stride, steps, d_cyc, d_sink);   // warm the L2 where it can be warm
            global_chase<<<ntab, 32 * w>>>(d_tabs, stride, steps, d_cyc, d_sink);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(h.data(), d_cyc, 8 * ntab, cudaMemcpyDeviceToHost));
            h.resize(ntab); printf(" %8.1f", avg(h) / steps); h.resize(4096);
        }
        printf("\n");
        CK(cudaFree(d_tabs));
    }
    {
        const int n = 2000;
        cta_barrier_cost<<<148, 512>>>(n, d_cyc); CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h.data(), d_cyc, 8, cudaMemcpyDeviceToHost));
        printf("barrier of one CTA of 16 warps: %.0f cycles", (double)h[0] / n);
        CK(cudaFuncSetAttribute(cluster_barrier_cost, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(9 * CL); cfg.blockDim = dim3(128);
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, cluster_barrier_cost, n, d_cyc);
        if (e == cudaSuccess) {
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(h.data(), d_cyc, 8, cudaMemcpyDeviceToHost));
            printf(";  barrier of a 16-CTA cluster (4 warps each): %.0f cycles\n", (double)h[0] / n);
        } else printf(";  cluster barrier: %s\n", cudaGetErrorString(e));
    }
    return 0;
}
