// micro: shared-memory atomics and partition traffic for the slab back end
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <random>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("ERR %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ unsigned long long gt() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

// MODE 0: smem atomicOr w/ return, random over window W words; 1: smem RED or (no return); 2: smem atomicAdd on small hist H cells
// 3: LDS of 8 words at random group; 4: STS 32 B record at random slot
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(const unsigned* __restrict__ keys, int per, int W, long long* out_clk, unsigned* sink) {
    extern __shared__ unsigned sm[];
    unsigned* s_key = sm;            // per
    unsigned* win = sm + per;        // W words
    const int tid = threadIdx.x, T = blockDim.x;
    for (int j = tid; j < per; j += T) s_key[j] = keys[(size_t)blockIdx.x * per + j];
    for (int j = tid; j < W; j += T) win[j] = 0;
    __syncthreads();
    long long t0 = clock64();
    unsigned acc = 0;
    for (int j0 = tid; j0 < per; j0 += 4 * T) {
        unsigned k4[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { int j = j0 + u * T; k4[u] = j < per ? s_key[j] : 0xffffffffu; }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (k4[u] == 0xffffffffu) continue;
            const unsigned key = k4[u] % ((unsigned)W * 32u);
            if (MODE == 0) acc += atomicOr(&win[key >> 5], 1u << (key & 31)) & (1u << (key & 31));
            if (MODE == 1) atomicOr(&win[key >> 5], 1u << (key & 31));
            if (MODE == 2) atomicAdd(&win[key % 600u], 1u);
            if (MODE == 3) { const uint4* p = (const uint4*)&win[(key >> 8) * 8]; uint4 a = p[0], b = p[1]; acc += __popc(a.x) + __popc(a.y)+__popc(a.z)+__popc(a.w)+__popc(b.x)+__popc(b.y)+__popc(b.z)+__popc(b.w); }
            if (MODE == 4) { uint4* p = (uint4*)&win[((key >> 5) % (unsigned)(W / 8)) * 8]; p[0] = make_uint4(key, 1, 2, 3); p[1] = make_uint4(4, 5, 6, 7); }
        }
    }
    __syncthreads();
    long long t1 = clock64();
    if (tid == 0) out_clk[blockIdx.x] = t1 - t0;
    if (acc == 0x12345678u) sink[0] = acc + win[tid];
}

// partition traffic: each CTA writes `per` 20-byte items (SoA key 4 + pt 16) in 148 runs to 148 buckets, then (separate kernel) reads its bucket
__global__ void __launch_bounds__(512, 1) k_scatter(const float4* __restrict__ pts, int per, unsigned* bkey, float4* bpts, int cap) {
    extern __shared__ unsigned sm[];
    float4* s_pts = (float4*)sm;
    const int tid = threadIdx.x, T = blockDim.x, b = blockIdx.x, G = gridDim.x;
    for (int j = tid; j < per; j += T) s_pts[j] = pts[(size_t)b * per + j];
    __syncthreads();
    const int run = per / G;   // items per destination
    // item s (sorted order) -> dest d = s / run, offset in bucket = b*run + s%run ; source slot = permuted
    for (int s = tid; s < run * G; s += T) {
        const int d = s / run, o = s - d * run;
        const int src = (int)(((unsigned)s * 2654435761u) % (unsigned)per);
        const float4 q = s_pts[src];
        bkey[(size_t)d * cap + b * run + o] = (unsigned)src;
        bpts[(size_t)d * cap + b * run + o] = q;
    }
}
__global__ void __launch_bounds__(512, 1) k_gather(const unsigned* bkey, const float4* bpts, int cap, int cnt, float* out) {
    const int tid = threadIdx.x, T = blockDim.x, b = blockIdx.x;
    float acc = 0;
    for (int j = tid; j < cnt; j += T) { acc += bpts[(size_t)b * cap + j].x + (float)bkey[(size_t)b * cap + j]; }
    if (acc == 1.2345f) out[0] = acc;
}

int main() {
    const int G = 148, per = 6784, W = 32768;   // 128 KB window
    std::vector<unsigned> h((size_t)G * per);
    std::mt19937 rng(1);
    for (auto& v : h) v = rng();
    unsigned* d; CK(cudaMalloc(&d, h.size() * 4)); CK(cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    long long* clk; CK(cudaMalloc(&clk, G * 8)); unsigned* sink; CK(cudaMalloc(&sink, 4096));
    const size_t smem = (size_t)(per + W) * 4;
    auto run = [&](auto kern, const char* name) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        for (int r = 0; r < 3; ++r) kern<<<G, 512, smem>>>(d, per, W, clk, sink);
        CK(cudaDeviceSynchronize());
        std::vector<long long> c(G); CK(cudaMemcpy(c.data(), clk, G * 8, cudaMemcpyDeviceToHost));
        long long mx = 0, sum = 0; for (auto v : c) { mx = v > mx ? v : mx; sum += v; }
        printf("%-40s  max %lld clk  avg %lld clk  (%.2f us @1.9GHz)  per point %.2f clk\n", name, mx, sum / G, mx / 1900.0, (double)mx / per);
    };
    run(k<0>, "smem atomicOr return, random 128KB");
    run(k<1>, "smem atomicOr no return");
    run(k<2>, "smem atomicAdd hist 600 cells");
    run(k<3>, "LDS 2x128 random group + popc");
    run(k<4>, "STS 32B record random slot");
    // partition traffic
    float4* pts; CK(cudaMalloc(&pts, (size_t)G * per * 16)); CK(cudaMemset(pts, 0, (size_t)G * per * 16));
    const int cap = 8192; unsigned* bkey; float4* bpts; CK(cudaMalloc(&bkey, (size_t)G * cap * 4)); CK(cudaMalloc(&bpts, (size_t)G * cap * 16));
    float* outf; CK(cudaMalloc(&outf, 4));
    CK(cudaFuncSetAttribute(k_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, per * 16));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int r = 0; r < 5; ++r) {
        CK(cudaEventRecord(e0)); k_scatter<<<G, 512, per * 16>>>(pts, per, bkey, bpts, cap); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); printf("scatter (load 16B/pt + write 20B/pt in 148 runs): %.1f us\n", ms * 1e3);
        CK(cudaEventRecord(e0)); k_gather<<<G, 512>>>(bkey, bpts, cap, (per / G) * G, outf); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1)); printf("gather bucket (read 20B/pt): %.1f us\n", ms * 1e3);
    }
    return 0;
}
