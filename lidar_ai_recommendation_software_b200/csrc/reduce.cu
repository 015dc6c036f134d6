// K1 — bounding box / moment reductions.
//
// Replaces the numpy reductions of the reference's preprocess stage:
//   np.min / np.max          utils/data_processing.py:143,207-208 ; app_simplified.py:80,116-117
//   np.mean / np.std axis 0  utils/data_processing.py:151-152     ; app_simplified.py:88-89
//
// One persistent launch: grid = k * SM count CTAs, grid-stride float4 / fp64 loads, warp shuffle
// tree, one partial per CTA, and the last CTA to finish (atomic ticket) folds the partials in a
// fixed order, so results are deterministic for a fixed launch shape.  HBM-bound: 16 B (F32X4)
// or 24 B (F64X3) per point, no writes.
#include "common.cuh"

namespace lidar {

constexpr int kRedThreads = 256;
constexpr int kRedMaxBlocks = 1024;

struct ReduceWs {
    double partial[kRedMaxBlocks][16];
    unsigned int ticket;
};

// The last CTA folds the per-CTA partials.  All 256 threads take part -- thread t folds channel t % 8 of the CTAs
// t / 8, t / 8 + 32, ... and the 32 partial results of a channel are then folded in a fixed order -- instead of
// eight threads walking up to 592 partials one dependent L2 load at a time (that tail was most of the launch on
// a 1 M-point cloud).  The order is fixed by the launch shape, so results stay deterministic.
template <class Op>
__device__ __forceinline__ void fold_partials(const ReduceWs* ws, int channels, double identity, Op op, double* out) {
    __shared__ double s_fold[32][8];
    const int ch = threadIdx.x & 7, row = threadIdx.x >> 3;       // kRedThreads == 256: 32 rows of 8 channels
    double v = identity;
    if (ch < channels)
        for (unsigned b = row; b < gridDim.x; b += 32) v = op(v, __ldcg(&ws->partial[b][ch]), ch);
    s_fold[row][ch] = v;
    __syncthreads();
    if (threadIdx.x < channels) {
        double r = s_fold[0][threadIdx.x];
        for (int k = 1; k < 32; ++k) r = op(r, s_fold[k][threadIdx.x], (int)threadIdx.x);
        out[threadIdx.x] = r;
    }
}
static_assert(kRedThreads == 256, "fold_partials assumes 32 rows of 8 channels");

template <class Loader>
__global__ void __launch_bounds__(kRedThreads)
bbox_kernel(Loader L, int64_t n, double* __restrict__ out8, ReduceWs* __restrict__ ws) {
    double mn[4] = {INFINITY, INFINITY, INFINITY, INFINITY};
    double mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    auto fold = [&](const Pt& p) {
        mn[0] = fmin(mn[0], p.x); mx[0] = fmax(mx[0], p.x);
        mn[1] = fmin(mn[1], p.y); mx[1] = fmax(mx[1], p.y);
        mn[2] = fmin(mn[2], p.z); mx[2] = fmax(mx[2], p.z);
        mn[3] = fmin(mn[3], p.w); mx[3] = fmax(mx[3], p.w);
    };
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n; i += 4 * stride) {      // four points (12 loads of the fp64 layout) in flight
        const Pt a = L.load(i), b = L.load(i + stride), c = L.load(i + 2 * stride), d = L.load(i + 3 * stride);
        fold(a); fold(b); fold(c); fold(d);
    }
    for (; i < n; i += stride) fold(L.load(i));
    __shared__ double s_mn[kRedThreads / 32][4];
    __shared__ double s_mx[kRedThreads / 32][4];
    __shared__ bool s_last;
    const int warp = threadIdx.x >> 5;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        mn[c] = warp_min(mn[c]);
        mx[c] = warp_max(mx[c]);
    }
    if (lane_id() == 0) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            s_mn[warp][c] = mn[c];
            s_mx[warp][c] = mx[c];
        }
    }
    __syncthreads();
    if (threadIdx.x < 8) {
        const int c = threadIdx.x & 3;
        const bool is_max = threadIdx.x >= 4;
        double v = is_max ? -INFINITY : INFINITY;
        for (int w = 0; w < kRedThreads / 32; ++w) v = is_max ? fmax(v, s_mx[w][c]) : fmin(v, s_mn[w][c]);
        ws->partial[blockIdx.x][threadIdx.x] = v;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = atomicAdd(&ws->ticket, 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    {
        const double ident = (threadIdx.x & 7) >= 4 ? -INFINITY : INFINITY;
        fold_partials(ws, 8, ident, [](double a, double q, int ch) { return ch >= 4 ? fmax(a, q) : fmin(a, q); }, out8);
    }
    if (threadIdx.x == 0) ws->ticket = 0u;
}

// float4 frames: min / max are exact in fp32 (they select, never round), so the loop stays in fp32 with four
// independent 16-byte loads in flight per thread and widens once at the end.  (The generic kernel converts
// every component to fp64 first — four quarter-rate F2F per point — and keeps one load in flight: 0.60 of the
// HBM peak on a 50 M-point scan.)
__global__ void __launch_bounds__(kRedThreads)
bbox_f32x4_kernel(LoadF32x4 L, int64_t n, double* __restrict__ out8, ReduceWs* __restrict__ ws) {
    float mn[4] = {INFINITY, INFINITY, INFINITY, INFINITY};
    float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    auto fold = [&](const float4& v) {
        mn[0] = fminf(mn[0], v.x); mx[0] = fmaxf(mx[0], v.x);
        mn[1] = fminf(mn[1], v.y); mx[1] = fmaxf(mx[1], v.y);
        mn[2] = fminf(mn[2], v.z); mx[2] = fmaxf(mx[2], v.z);
        mn[3] = fminf(mn[3], v.w); mx[3] = fmaxf(mx[3], v.w);
    };
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n; i += 4 * stride) {
        const float4 a = L.raw(i), b = L.raw(i + stride), c = L.raw(i + 2 * stride), d = L.raw(i + 3 * stride);
        fold(a); fold(b); fold(c); fold(d);
    }
    for (; i < n; i += stride) fold(L.raw(i));
    __shared__ float s_v[kRedThreads / 32][8];
    __shared__ bool s_last;
    const int warp = threadIdx.x >> 5;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o));
            mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
        }
    }
    if (lane_id() == 0) {
#pragma unroll
        for (int c = 0; c < 4; ++c) { s_v[warp][c] = mn[c]; s_v[warp][4 + c] = mx[c]; }
    }
    __syncthreads();
    if (threadIdx.x < 8) {
        const bool is_max = threadIdx.x >= 4;
        float v = is_max ? -INFINITY : INFINITY;
        for (int w = 0; w < kRedThreads / 32; ++w) v = is_max ? fmaxf(v, s_v[w][threadIdx.x]) : fminf(v, s_v[w][threadIdx.x]);
        ws->partial[blockIdx.x][threadIdx.x] = (double)v;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&ws->ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    {
        const double ident = (threadIdx.x & 7) >= 4 ? -INFINITY : INFINITY;
        fold_partials(ws, 8, ident, [](double a, double q, int ch) { return ch >= 4 ? fmax(a, q) : fmin(a, q); }, out8);
    }
    if (threadIdx.x == 0) ws->ticket = 0u;
}

// Σ (p - c) and Σ (p - c)^2 per axis, fp64, c = caller-supplied centre (0 for the mean pass).
// out6 = {Σdx, Σdy, Σdz, Σdx², Σdy², Σdz²}
template <class Loader>
__global__ void __launch_bounds__(kRedThreads)
moments_kernel(Loader L, int64_t n, double cx, double cy, double cz, double* __restrict__ out6,
               ReduceWs* __restrict__ ws, const long long* __restrict__ d_n = nullptr,
               const double* __restrict__ d_center3 = nullptr) {
    if (d_n) n = *d_n;                        // chained preprocess: count and centre come from the previous stage
    if (d_center3) { cx = d_center3[0]; cy = d_center3[1]; cz = d_center3[2]; }
    double s[6] = {0, 0, 0, 0, 0, 0};
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    auto fold = [&](const Pt& p) {
        const double dx = p.x - cx, dy = p.y - cy, dz = p.z - cz;
        s[0] += dx; s[1] += dy; s[2] += dz;
        s[3] += __dmul_rn(dx, dx); s[4] += __dmul_rn(dy, dy); s[5] += __dmul_rn(dz, dz);
    };
    // the loads of four points are issued together; the adds stay in index order (same sums as the plain loop)
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n; i += 4 * stride) {
        const Pt a = L.load(i), b = L.load(i + stride), c = L.load(i + 2 * stride), d = L.load(i + 3 * stride);
        fold(a); fold(b); fold(c); fold(d);
    }
    for (; i < n; i += stride) fold(L.load(i));
    __shared__ double s_p[kRedThreads / 32][6];
    __shared__ bool s_last;
    const int warp = threadIdx.x >> 5;
#pragma unroll
    for (int c = 0; c < 6; ++c) s[c] = warp_sum(s[c]);
    if (lane_id() == 0) {
#pragma unroll
        for (int c = 0; c < 6; ++c) s_p[warp][c] = s[c];
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        double v = 0;
        for (int w = 0; w < kRedThreads / 32; ++w) v += s_p[w][threadIdx.x];
        ws->partial[blockIdx.x][threadIdx.x] = v;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = atomicAdd(&ws->ticket, 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    fold_partials(ws, 6, 0.0, [](double a, double q, int) { return a + q; }, out6);
    if (threadIdx.x == 0) ws->ticket = 0u;
}

static int reduce_grid(int64_t n) {
    int64_t want = (n + kRedThreads * 4 - 1) / (kRedThreads * 4);
    int64_t cap = (int64_t)sm_count() * 4;
    if (cap > kRedMaxBlocks) cap = kRedMaxBlocks;
    if (want < 1) want = 1;
    return (int)(want < cap ? want : cap);
}

// lidar_moments with the row count and the centre in device memory (lidar_preprocess_front); the grid is sized by
// the capacity, so the fold order is fixed by `cap` alone
int moments_f64x3_dev(const double* d_points, int64_t cap, const long long* d_n, const double* d_center3,
                      double* d_out6, void* d_reduce_ws, cudaStream_t st) {
    ReduceWs* ws = static_cast<ReduceWs*>(d_reduce_ws);
    moments_kernel<<<reduce_grid(cap), kRedThreads, 0, st>>>(LoadF64x3{d_points}, cap, 0.0, 0.0, 0.0, d_out6, ws, d_n,
                                                             d_center3);
    LIDAR_CHECK_LAUNCH();
    return LIDAR_OK;
}

// utils/visualization.py:50-54: centroid = np.mean(points, axis=0); distances = sqrt(sum((p - centroid)^2, axis=1)).
// The centroid comes from the device moments (sum / n, evaluated on the device from d_sum6 / n); the per-point part
// follows numpy's expression: three squared differences added left to right, one correctly rounded sqrt.
__global__ void __launch_bounds__(kRedThreads)
centroid_distance_kernel(const double* __restrict__ pts, int64_t n, const double* __restrict__ sum6, double* __restrict__ out) {
    const double inv_n = (double)n;
    const double cx = __ddiv_rn(sum6[0], inv_n), cy = __ddiv_rn(sum6[1], inv_n), cz = __ddiv_rn(sum6[2], inv_n);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double dx = __dsub_rn(__ldg(pts + 3 * i), cx), dy = __dsub_rn(__ldg(pts + 3 * i + 1), cy);
        const double dz = __dsub_rn(__ldg(pts + 3 * i + 2), cz);
        out[i] = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
    }
}

}  // namespace lidar

using namespace lidar;

extern "C" {

int lidar_centroid_distances(const double* d_points, int64_t n, double* d_out, void* d_ws, size_t ws_bytes, void* stream) {
    LIDAR_REQUIRE(n >= 0 && (n == 0 || (d_points && d_out)), LIDAR_ERR_INVALID, "lidar_centroid_distances: bad argument");
    LIDAR_REQUIRE(d_ws && ws_bytes >= ws_align(sizeof(ReduceWs)) + 64, LIDAR_ERR_WORKSPACE,
                  "lidar_centroid_distances: workspace too small");
    if (n == 0) return LIDAR_OK;
    cudaStream_t st = as_stream(stream);
    double* sum6 = reinterpret_cast<double*>(static_cast<char*>(d_ws) + ws_align(sizeof(ReduceWs)));
    const double zero[3] = {0.0, 0.0, 0.0};
    const int rc = lidar_moments(d_points, LIDAR_FMT_F64X3, n, zero, sum6, d_ws, ws_align(sizeof(ReduceWs)), stream);
    if (rc != LIDAR_OK) return rc;
    centroid_distance_kernel<<<reduce_grid(n), kRedThreads, 0, st>>>(d_points, n, sum6, d_out);
    LIDAR_CHECK_LAUNCH();
    return LIDAR_OK;
}

size_t lidar_reduce_workspace_bytes(void) { return ws_align(sizeof(ReduceWs)) + 256; }

int lidar_bbox(const void* d_points, int fmt, int64_t n, double* d_out8, void* d_ws, size_t ws_bytes,
               void* stream) {
    LIDAR_REQUIRE(n >= 0, LIDAR_ERR_INVALID, "lidar_bbox: n < 0");
    LIDAR_REQUIRE(d_out8 != nullptr, LIDAR_ERR_INVALID, "lidar_bbox: d_out8 is NULL");
    LIDAR_REQUIRE(d_ws && ws_bytes >= sizeof(ReduceWs), LIDAR_ERR_WORKSPACE,
                  "lidar_bbox: workspace too small (%zu < %zu)", ws_bytes, sizeof(ReduceWs));
    LIDAR_REQUIRE(n == 0 || d_points != nullptr, LIDAR_ERR_INVALID, "lidar_bbox: d_points is NULL");
    cudaStream_t st = as_stream(stream);
    ReduceWs* ws = static_cast<ReduceWs*>(d_ws);
    LIDAR_CUDA_TRY(cudaMemsetAsync(&ws->ticket, 0, sizeof(unsigned), st));
    const int grid = reduce_grid(n);
    if (fmt == LIDAR_FMT_F32X4) {
        bbox_f32x4_kernel<<<grid, kRedThreads, 0, st>>>(LoadF32x4{static_cast<const float4*>(d_points)}, n, d_out8, ws);
    } else if (fmt == LIDAR_FMT_F64X3) {
        bbox_kernel<<<grid, kRedThreads, 0, st>>>(LoadF64x3{static_cast<const double*>(d_points)}, n, d_out8, ws);
    } else {
        LIDAR_REQUIRE(false, LIDAR_ERR_INVALID, "lidar_bbox: unknown point format %d", fmt);
    }
    LIDAR_CHECK_LAUNCH();
    return LIDAR_OK;
}

int lidar_moments(const void* d_points, int fmt, int64_t n, const double* h_center3, double* d_out6,
                  void* d_ws, size_t ws_bytes, void* stream) {
    LIDAR_REQUIRE(n >= 0, LIDAR_ERR_INVALID, "lidar_moments: n < 0");
    LIDAR_REQUIRE(d_out6 != nullptr && h_center3 != nullptr, LIDAR_ERR_INVALID, "lidar_moments: NULL argument");
    LIDAR_REQUIRE(d_ws && ws_bytes >= sizeof(ReduceWs), LIDAR_ERR_WORKSPACE, "lidar_moments: workspace too small");
    cudaStream_t st = as_stream(stream);
    ReduceWs* ws = static_cast<ReduceWs*>(d_ws);
    LIDAR_CUDA_TRY(cudaMemsetAsync(&ws->ticket, 0, sizeof(unsigned), st));
    const int grid = reduce_grid(n);
    if (fmt == LIDAR_FMT_F32X4) {
        moments_kernel<<<grid, kRedThreads, 0, st>>>(LoadF32x4{static_cast<const float4*>(d_points)}, n,
                                                     h_center3[0], h_center3[1], h_center3[2], d_out6, ws);
    } else if (fmt == LIDAR_FMT_F64X3) {
        moments_kernel<<<grid, kRedThreads, 0, st>>>(LoadF64x3{static_cast<const double*>(d_points)}, n,
                                                     h_center3[0], h_center3[1], h_center3[2], d_out6, ws);
    } else {
        LIDAR_REQUIRE(false, LIDAR_ERR_INVALID, "lidar_moments: unknown point format %d", fmt);
    }
    LIDAR_CHECK_LAUNCH();
    return LIDAR_OK;
}

}  // extern "C"
