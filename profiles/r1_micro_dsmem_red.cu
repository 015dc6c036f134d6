// micro: throughput of remote shared-memory reductions (red.shared::cluster.add.u32) inside a thread-block cluster
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <cstdio>
#include <vector>
namespace cg = cooperative_groups;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("ERR %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ unsigned hash32(unsigned x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

template <int MODE>   // 0: remote red spread over the cluster, 1: local smem red only, 2: global RED (L2) to same-size table
__global__ void __launch_bounds__(512, 1) k(int cells_per_cta, int per_thread, unsigned* gtab, long long* clk) {
    extern __shared__ unsigned sm[];
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned C = cluster.num_blocks(), rank = cluster.block_rank();
    for (int i = threadIdx.x; i < cells_per_cta; i += blockDim.x) sm[i] = 0;
    cluster.sync();
    const unsigned base = (unsigned)__cvta_generic_to_shared(sm);
    unsigned seed = (blockIdx.x * 512u + threadIdx.x) * 2654435761u + 12345u;
    const unsigned total_cells = C * (unsigned)cells_per_cta;
    long long t0 = clock64();
    for (int i = 0; i < per_thread; ++i) {
        seed = hash32(seed + i);
        const unsigned cell = seed % total_cells;
        if (MODE == 0) {
            const unsigned owner = cell / (unsigned)cells_per_cta, local = cell - owner * (unsigned)cells_per_cta;
            unsigned raddr;
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(base + local * 4u), "r"(owner));
            asm volatile("red.relaxed.cluster.shared::cluster.add.u32 [%0], %1;" ::"r"(raddr), "r"(1u) : "memory");
        } else if (MODE == 1) {
            atomicAdd(&sm[cell % (unsigned)cells_per_cta], 1u);
        } else {
            asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(gtab + (size_t)(blockIdx.x / C) * total_cells + cell), "r"(1u) : "memory");
        }
    }
    cluster.sync();
    long long t1 = clock64();
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
    if (sm[threadIdx.x] == 0xdeadbeef) gtab[0] = 1;
}

template <int MODE>
void run(const char* name, int csize, int cells_per_cta, int per_thread, unsigned* gtab, long long* clk) {
    const int G = (148 / csize) * csize;
    cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(G); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = (size_t)cells_per_cta * 4; cfg.stream = 0;
    cudaLaunchAttribute attr[1]; attr[0].id = cudaLaunchAttributeClusterDimension; attr[0].val.clusterDim.x = csize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    CK(cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, cells_per_cta * 4));
    if (csize > 8) CK(cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e9;
    for (int r = 0; r < 3; ++r) {
        CK(cudaEventRecord(e0));
        cudaError_t e = cudaLaunchKernelEx(&cfg, k<MODE>, cells_per_cta, per_thread, gtab, clk);
        if (e != cudaSuccess) { printf("%-40s launch failed: %s\n", name, cudaGetErrorString(e)); cudaGetLastError(); return; }
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = ms < best ? ms : best;
    }
    std::vector<long long> c(G); CK(cudaMemcpy(c.data(), clk, G * 8, cudaMemcpyDeviceToHost));
    long long mx = 0; for (auto v : c) mx = v > mx ? v : mx;
    const double ops = (double)G * 512 * per_thread;
    printf("%-40s cluster %2d ctas %3d  %.3f ms  %.1f Gops/s  %.3f ops/clk/SM (loop clk %lld)\n", name, csize, G, best, ops / best / 1e6, (double)512 * per_thread / mx, mx);
}
int main() {
    unsigned* gtab; CK(cudaMalloc(&gtab, 64ull << 20)); CK(cudaMemset(gtab, 0, 64ull << 20));
    long long* clk; CK(cudaMalloc(&clk, 148 * 8));
    const int cells = 50000, per = 2048;
    run<0>("remote red, cluster-wide random", 8, cells, per, gtab, clk);
    run<0>("remote red, cluster-wide random", 4, cells, per, gtab, clk);
    run<0>("remote red, cluster-wide random", 2, cells, per, gtab, clk);
    run<0>("remote red (all local, cluster 1)", 1, cells, per, gtab, clk);
    run<0>("remote red, cluster-wide random", 16, 30000, per, gtab, clk);
    run<1>("local atomicAdd smem", 8, cells, per, gtab, clk);
    run<2>("global RED same table size", 8, cells, per, gtab, clk);
    return 0;
}
