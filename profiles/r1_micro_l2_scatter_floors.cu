// micro-benchmarks: floors of the scattered accesses of the frame pipeline (1 M points, 148 x 512 threads)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <random>
#include <cmath>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("ERR %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
struct __align__(32) W8 { unsigned w[8]; };
__device__ __forceinline__ W8 ld256(const void* p) { W8 v; asm volatile("ld.global.cg.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(v.w[0]),"=r"(v.w[1]),"=r"(v.w[2]),"=r"(v.w[3]),"=r"(v.w[4]),"=r"(v.w[5]),"=r"(v.w[6]),"=r"(v.w[7]) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st256(void* p, const W8& v) { asm volatile("st.global.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" :: "l"(p),"r"(v.w[0]),"r"(v.w[1]),"r"(v.w[2]),"r"(v.w[3]),"r"(v.w[4]),"r"(v.w[5]),"r"(v.w[6]),"r"(v.w[7]) : "memory"); }

template <int MODE, int U>
__global__ void __launch_bounds__(512, 1) k(const int* __restrict__ key, const int* __restrict__ cell, int n, unsigned* bitmap, int* grid, W8* rec, int* out) {
    const int stride = gridDim.x * blockDim.x;
    unsigned acc = 0;
    for (int i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += U * stride) {
        int kk[U], cc[U]; unsigned old[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { int i = i0 + u * stride; kk[u] = i < n ? key[i] : -1; if (MODE == 2 || MODE == 3 || MODE >= 8) cc[u] = i < n ? cell[i] : -1; }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (kk[u] < 0) continue;
            const unsigned g = (unsigned)kk[u] / 224u, b = (unsigned)kk[u] - g * 224u;
            unsigned* w = bitmap + (size_t)g * 8 + 1 + (b >> 5);
            if (MODE == 0 || MODE == 3) old[u] = atomicOr(w, 1u << (b & 31));           // ATOM with return
            if (MODE == 1) atomicOr(w, 1u << (b & 31));                                  // RED
            if (MODE == 8) { old[u] = *(volatile unsigned*)w; }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (kk[u] < 0) continue;
            if (MODE == 2 || MODE == 3 || MODE == 8) atomicAdd(grid + cc[u], 1);
            if (MODE == 9) atomicAdd(grid + (blockIdx.x & 7) * 65536 + cc[u], 1);
            if (MODE == 10) atomicAdd(grid + (blockIdx.x & 1) * 65536 + cc[u], 1);
            if (MODE == 11) atomicAdd(grid + cc[u] * 8, 1);   // one counter per 32-byte sector
            if (MODE == 0 || MODE == 3 || MODE == 8) acc += old[u];
        }
        if (MODE == 4 || MODE == 6) {
            W8 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) if (kk[u] >= 0) v[u] = ld256(bitmap + (size_t)((unsigned)kk[u] / 224u) * 8);
#pragma unroll
            for (int u = 0; u < U; ++u) if (kk[u] >= 0) {
                acc += v[u].w[0] + v[u].w[3] + v[u].w[7];
                if (MODE == 6) { W8 r = v[u]; r.w[1] = kk[u]; st256(rec + (i0 + u * stride) * 7919ull % n, r); }
            }
        }
        if (MODE == 5) {
#pragma unroll
            for (int u = 0; u < U; ++u) if (kk[u] >= 0) { W8 r; for (int q = 0; q < 8; ++q) r.w[q] = kk[u] + q; st256(rec + ((unsigned long long)(i0 + u * stride) * 7919ull) % n, r); }
        }
        if (MODE == 7) {   // two LDG.128 instead of one 256
#pragma unroll
            for (int u = 0; u < U; ++u) if (kk[u] >= 0) { const uint4* p = (const uint4*)(bitmap + (size_t)((unsigned)kk[u] / 224u) * 8); uint4 a = __ldcg(p), c = __ldcg(p + 1); acc += a.x + c.w; }
        }
    }
    if (acc == 0x12345678u) out[0] = 1;
}

template <int MODE, int U>
float run(const char* name, const int* key, const int* cell, int n, unsigned* bitmap, size_t bm_bytes, int* grid, W8* rec, int* out, int threads, int ctas) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9, sum = 0;
    for (int rep = 0; rep < 12; ++rep) {
        CK(cudaMemsetAsync(bitmap, 0, bm_bytes));
        CK(cudaMemsetAsync(grid, 0, 8 * 256 * 256 * 4));
        cudaEventRecord(e0);
        k<MODE, U><<<ctas, threads>>>(key, cell, n, bitmap, grid, rec, out);
        cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep >= 2) { best = fminf(best, ms); sum += ms; }
    }
    printf("%-44s U=%d thr=%d ctas=%d  best %.1f us  avg %.1f us\n", name, U, threads, ctas, best * 1e3, sum / 10 * 1e3);
    return best;
}

int main() {
    const int n = 1000000;
    std::mt19937_64 rng(1);
    std::vector<int> key(n), cell(n), cellu(n);
    // crowd-like: 40 % ground uniform over 2000x2000x(z small), 60 % people in 1500 clusters
    std::uniform_real_distribution<double> U01(0, 1);
    std::normal_distribution<double> N01(0, 1);
    std::vector<double> cx(1500), cy(1500);
    for (int p = 0; p < 1500; ++p) { cx[p] = -45 + 90 * U01(rng); cy[p] = -45 + 90 * U01(rng); }
    for (int i = 0; i < n; ++i) {
        double x, y, z;
        if (U01(rng) < 0.4) { x = -50 + 100 * U01(rng); y = -50 + 100 * U01(rng); z = 0.1 * U01(rng); }
        else { int p = (int)(U01(rng) * 1500); x = cx[p] + 0.12 * N01(rng); y = cy[p] + 0.12 * N01(rng); z = 0.1 + 1.7 * U01(rng); }
        int ix = (int)((x + 50) / 0.05), iy = (int)((y + 50) / 0.05), iz = (int)((z + 0.15) / 0.05);
        ix = std::min(std::max(ix, 0), 1999); iy = std::min(std::max(iy, 0), 1999); iz = std::min(std::max(iz, 0), 38);
        key[i] = (ix * 2000 + iy) * 39 + iz;
        int bx = (int)((x + 51) / 0.5), by = (int)((y + 51) / 0.5);
        cell[i] = bx * 204 + by;
        cellu[i] = (int)(U01(rng) * 204 * 204);
    }
    int *dkey, *dcell, *dcellu, *grid, *out; unsigned* bitmap; W8* rec;
    size_t bm = (size_t)(156000000 / 224 + 1024) * 32;
    CK(cudaMalloc(&dkey, n * 4)); CK(cudaMalloc(&dcell, n * 4)); CK(cudaMalloc(&dcellu, n * 4)); CK(cudaMalloc(&grid, 8 * 256 * 256 * 4));
    CK(cudaMalloc(&out, 4)); CK(cudaMalloc(&bitmap, bm)); CK(cudaMalloc(&rec, (size_t)n * 32));
    CK(cudaMemcpy(dkey, key.data(), n * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dcell, cell.data(), n * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dcellu, cellu.data(), n * 4, cudaMemcpyHostToDevice));
    for (int cfg = 0; cfg < 1; ++cfg) {
        const int threads = cfg == 2 ? 256 : 512, ctas = cfg == 0 ? 148 : (cfg == 1 ? 296 : 148 * 8);
        printf("---- threads %d ctas %d\n", threads, ctas);
        run<0, 1>("atomicOr (return) random bitmap", dkey, dcell, n, bitmap, bm, grid, rec, out, threads, ctas);
        run<0, 4>("atomicOr (return) random bitmap", dkey, dcell, n, bitmap, bm, grid, rec, out, threads, ctas);
        run<1, 4>("RED.OR random bitmap", dkey, dcell, n, bitmap, bm, grid, rec, out, threads, ctas);
        run<2, 4>("RED.ADD grid crowd", dkey, dcell, n, bitmap, bm, grid, rec, out, threads, ctas);
        run<2, 4>("RED.ADD grid uniform", dkey, dcellu, n, bitmap, bm, grid, rec, out, threads, ctas);
        run<3, 4>("atomicOr + RED.ADD crowd", dkey, dcell, n, bitmap, bm, grid, rec, out, threads, ctas);
        run<3, 4>("atomicOr + RED.ADD uniform", dkey, dcellu, n, bitmap, bm, grid, rec, out, threads, ctas);
        run<9, 4>("RED.ADD grid crowd, 8 replicas", dkey, dcell, n, bitmap, bm, grid, rec, out, threads, ctas);
        run<10, 4>("RED.ADD grid crowd, 2 replicas", dkey, dcell, n, bitmap, bm, grid, rec, out, threads, ctas);
        run<11, 4>("RED.ADD grid crowd, 1 counter / sector", dkey, dcell, n, bitmap, bm, grid, rec, out, threads, ctas);
        run<8, 4>("plain load + RED.ADD crowd", dkey, dcell, n, bitmap, bm, grid, rec, out, threads, ctas);
        run<4, 4>("ld256 random group", dkey, dcell, n, bitmap, bm, grid, rec, out, threads, ctas);
        run<7, 4>("2x ld128 random group", dkey, dcell, n, bitmap, bm, grid, rec, out, threads, ctas);
        run<5, 4>("st256 scattered record", dkey, dcell, n, bitmap, bm, grid, rec, out, threads, ctas);
        run<6, 4>("ld256 + st256", dkey, dcell, n, bitmap, bm, grid, rec, out, threads, ctas);
        run<6, 1>("ld256 + st256", dkey, dcell, n, bitmap, bm, grid, rec, out, threads, ctas);
    }
    return 0;
}
