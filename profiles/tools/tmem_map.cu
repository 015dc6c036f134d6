// prints the (row, col) held by each register of tcgen05.ld.16x256b.x1 / .x2 for warp 0 (TMEM lanes 0..31)
#include <cstdio>
#include <cuda_runtime.h>
__device__ unsigned s_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__global__ void k(unsigned* out) {
    __shared__ unsigned slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_addr(&slot)), "r"(64) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned base = slot;
    // every thread writes its row: value = row * 1000 + col, 32 columns (two 32x32b.x16 stores)
    const unsigned lane_addr = base + ((unsigned)(warp * 32) << 16);
    for (int h = 0; h < 2; ++h) {
        unsigned v[16];
        for (int i = 0; i < 16; ++i) v[i] = (unsigned)((warp * 32 + lane) * 1000 + h * 16 + i);
        asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                     ::"r"(lane_addr + h * 16), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                       "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp == 1) {   // warp 1 reads ITS lanes (32..63): first half (lanes 32..47) and second half (48..63), columns 8..23
        for (int h = 0; h < 2; ++h) {
            unsigned r[8];
            const unsigned a = base + ((unsigned)(32 + 16 * h) << 16) + 8;
            asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(a) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int i = 0; i < 8; ++i) out[(h * 32 + lane) * 8 + i] = r[i];
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(64) : "memory");
}
int main() {
    unsigned* d; cudaMalloc(&d, 2 * 32 * 8 * 4);
    k<<<1, 128>>>(d);
    unsigned h[2 * 32 * 8];
    cudaError_t e = cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("err %d\n", (int)e);
    for (int hf = 0; hf < 2; ++hf)
        for (int t = 0; t < 32; ++t) {
            printf("half %d lane %2d:", hf, t);
            for (int i = 0; i < 8; ++i) printf(" r%d=(row %2u,col %2u)", i, h[(hf * 32 + t) * 8 + i] / 1000, h[(hf * 32 + t) * 8 + i] % 1000);
            printf("\n");
        }
    return 0;
}
