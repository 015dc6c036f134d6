// K13 — shared MLP (1x1 conv + folded-BN bias + ReLU, three layers) + max-pool over the k neighbours.
//
// NEW op (the reference has no network code).  Two implementations with the
// same C entry point:
//   * `simt`    fp32 CUDA-core kernel: reference implementation of the arithmetic, any layer widths.
//   * `tcgen05` (sa_mlp_tc.cu) 5th-gen tensor-core kernel for the 64-64-128 SSG configuration.
// Both fuse the grouping (gather + centre subtraction) into the load of the first layer, so the
// (B, 3+C, M, k) grouped tensor is never materialised, and the max-pool into the epilogue of the last
// layer (ReLU output >= 0, so an integer atomicMax on the fp32 bit pattern is an exact max).
#include "common.cuh"

namespace lidar {

constexpr int kMlpThreads = 256;
constexpr int kMlpRows = 64;   // rows (centre, neighbour) pairs per CTA tile

struct MlpDims {
    int c_in, c1, c2, c3;
};

// rows of the tile -> input features: either gathered (xyz[idx] - centre, feats[idx]) or read from a
// materialised grouped tensor (B, c_in, M, k)
struct MlpInput {
    const float* xyz;       // (B, N, 3)
    const float* feats;     // (B, C, N) or NULL
    const int* idx;         // (B, M, k)
    const float* new_xyz;   // (B, M, 3)
    const float* grouped;   // (B, c_in, M, k) or NULL
    int n, m, k;
};

__device__ __forceinline__ float mlp_in(const MlpInput& I, int b, int row /* m*k + j */, int c, int c_in) {
    if (I.grouped) return I.grouped[((size_t)b * c_in + c) * ((size_t)I.m * I.k) + row];
    const int mm = row / I.k;
    const int src = I.idx[(size_t)b * I.m * I.k + row];
    if (c < 3) return __fsub_rn(I.xyz[((size_t)b * I.n + src) * 3 + c], I.new_xyz[((size_t)b * I.m + mm) * 3 + c]);
    return I.feats[((size_t)b * (c_in - 3) + (c - 3)) * I.n + src];
}

// out[b][c3][m] must be zero-initialised (max of ReLU outputs)
__global__ void __launch_bounds__(kMlpThreads)
shared_mlp_simt_kernel(MlpInput I, MlpDims D, const float* __restrict__ W1, const float* __restrict__ B1,
                       const float* __restrict__ W2, const float* __restrict__ B2, const float* __restrict__ W3,
                       const float* __restrict__ B3, int rows_per_cloud, float* __restrict__ out) {
    extern __shared__ float smem[];
    // layout: W1[c1][c_in] | W2[c2][c1] | W3[c3][c2] | act0[rows][c_in] | act1[rows][c1+1] | act2[rows][c2+1]
    float* sW1 = smem;
    float* sW2 = sW1 + D.c1 * D.c_in;
    float* sW3 = sW2 + D.c2 * D.c1;
    float* a0 = sW3 + D.c3 * D.c2;
    float* a1 = a0 + kMlpRows * D.c_in;
    float* a2 = a1 + kMlpRows * (D.c1 + 1);
    for (int i = threadIdx.x; i < D.c1 * D.c_in; i += kMlpThreads) sW1[i] = W1[i];
    for (int i = threadIdx.x; i < D.c2 * D.c1; i += kMlpThreads) sW2[i] = W2[i];
    for (int i = threadIdx.x; i < D.c3 * D.c2; i += kMlpThreads) sW3[i] = W3[i];
    const int tiles_per_cloud = (rows_per_cloud + kMlpRows - 1) / kMlpRows;
    const int b = blockIdx.x / tiles_per_cloud;
    const int row0 = (blockIdx.x % tiles_per_cloud) * kMlpRows;
    const int rows = min(kMlpRows, rows_per_cloud - row0);
    for (int i = threadIdx.x; i < kMlpRows * D.c_in; i += kMlpThreads) {
        const int r = i / D.c_in, c = i % D.c_in;
        a0[i] = r < rows ? mlp_in(I, b, row0 + r, c, D.c_in) : 0.f;
    }
    __syncthreads();
    // layer 1
    for (int i = threadIdx.x; i < kMlpRows * D.c1; i += kMlpThreads) {
        const int r = i / D.c1, o = i % D.c1;
        float acc = B1[o];
        for (int c = 0; c < D.c_in; ++c) acc = fmaf(sW1[o * D.c_in + c], a0[r * D.c_in + c], acc);
        a1[r * (D.c1 + 1) + o] = fmaxf(acc, 0.f);
    }
    __syncthreads();
    // layer 2
    for (int i = threadIdx.x; i < kMlpRows * D.c2; i += kMlpThreads) {
        const int r = i % kMlpRows, o = i / kMlpRows;   // lanes walk rows: conflict-free a1 reads, W broadcast
        float acc = B2[o];
        const float* w = sW2 + o * D.c1;
        const float* a = a1 + r * (D.c1 + 1);
        for (int c = 0; c < D.c1; ++c) acc = fmaf(w[c], a[c], acc);
        a2[r * (D.c2 + 1) + o] = fmaxf(acc, 0.f);
    }
    __syncthreads();
    // layer 3 + max-pool
    for (int i = threadIdx.x; i < kMlpRows * D.c3; i += kMlpThreads) {
        const int r = i % kMlpRows, o = i / kMlpRows;
        if (r >= rows) continue;
        float acc = B3[o];
        const float* w = sW3 + o * D.c2;
        const float* a = a2 + r * (D.c2 + 1);
        for (int c = 0; c < D.c2; ++c) acc = fmaf(w[c], a[c], acc);
        acc = fmaxf(acc, 0.f);
        const int mm = (row0 + r) / I.k;
        atomicMax(reinterpret_cast<int*>(out + ((size_t)b * D.c3 + o) * I.m + mm), __float_as_int(acc));
    }
}

size_t mlp_simt_smem(const MlpDims& D) {
    return sizeof(float) * ((size_t)D.c1 * D.c_in + (size_t)D.c2 * D.c1 + (size_t)D.c3 * D.c2 +
                            (size_t)kMlpRows * D.c_in + (size_t)kMlpRows * (D.c1 + 1) + (size_t)kMlpRows * (D.c2 + 1));
}

// defined in sa_mlp_tc.cu
int launch_shared_mlp_tc(const float* xyz, const int* idx, const float* new_xyz, int b, int n, int m, int k,
                         const float* W1, const float* B1, const float* W2, const float* B2, const float* W3,
                         const float* B3, float* out, cudaStream_t st);

}  // namespace lidar

using namespace lidar;

extern "C" {

int lidar_shared_mlp_maxpool(const float* d_xyz, const float* d_feats, const int32_t* d_idx, const float* d_new_xyz,
                             const float* d_grouped, int b, int n, int m, int k, int c_in, int c1, int c2, int c3,
                             const float* d_w1, const float* d_b1, const float* d_w2, const float* d_b2,
                             const float* d_w3, const float* d_b3, float* d_out, int impl, void* stream) {
    LIDAR_REQUIRE(b >= 0 && m >= 1 && k >= 1 && c_in >= 1 && c1 >= 1 && c2 >= 1 && c3 >= 1, LIDAR_ERR_INVALID,
                  "lidar_shared_mlp_maxpool: bad sizes");
    LIDAR_REQUIRE(d_w1 && d_b1 && d_w2 && d_b2 && d_w3 && d_b3 && d_out, LIDAR_ERR_INVALID,
                  "lidar_shared_mlp_maxpool: NULL weights / output");
    LIDAR_REQUIRE(d_grouped || (d_xyz && d_idx && d_new_xyz && n >= 1 && (c_in == 3 || d_feats)), LIDAR_ERR_INVALID,
                  "lidar_shared_mlp_maxpool: need either a grouped tensor or xyz + idx + new_xyz (+ feats)");
    if (b == 0) return LIDAR_OK;
    cudaStream_t st = as_stream(stream);
    LIDAR_CUDA_TRY(cudaMemsetAsync(d_out, 0, sizeof(float) * (size_t)b * c3 * m, st));
    const bool tc_shape = !d_grouped && c_in == 3 && c1 == 64 && c2 == 64 && c3 == 128 && k == 32;
    if (impl == LIDAR_MLP_TCGEN05) {
        LIDAR_REQUIRE(tc_shape, LIDAR_ERR_INVALID,
                      "lidar_shared_mlp_maxpool: the tcgen05 kernel covers c_in=3, widths 64-64-128, k=32, fused gather");
        return launch_shared_mlp_tc(d_xyz, d_idx, d_new_xyz, b, n, m, k, d_w1, d_b1, d_w2, d_b2, d_w3, d_b3, d_out, st);
    }
    if (impl == LIDAR_MLP_AUTO && tc_shape)
        return launch_shared_mlp_tc(d_xyz, d_idx, d_new_xyz, b, n, m, k, d_w1, d_b1, d_w2, d_b2, d_w3, d_b3, d_out, st);
    MlpDims D{c_in, c1, c2, c3};
    const size_t smem = mlp_simt_smem(D);
    LIDAR_REQUIRE(smem <= smem_optin(), LIDAR_ERR_INVALID, "lidar_shared_mlp_maxpool: layer widths need %zu B of shared memory", smem);
    LIDAR_CUDA_TRY(cudaFuncSetAttribute(shared_mlp_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MlpInput I{d_xyz, d_feats, d_idx, d_new_xyz, d_grouped, n, m, k};
    const int rows_per_cloud = m * k;
    const int tiles = (rows_per_cloud + kMlpRows - 1) / kMlpRows;
    shared_mlp_simt_kernel<<<b * tiles, kMlpThreads, smem, st>>>(I, D, d_w1, d_b1, d_w2, d_b2, d_w3, d_b3,
                                                                 rows_per_cloud, d_out);
    LIDAR_CHECK_LAUNCH();
    return LIDAR_OK;
}

}  // extern "C"
