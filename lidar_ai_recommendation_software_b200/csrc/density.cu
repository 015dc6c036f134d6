// K6 — 2-D density histogram with numpy.histogramdd semantics, bit-exact integer counts.
//
// Replaces np.histogram2d at
//   utils/data_processing.py:316-319   (calculate_grid_density, np.arange edges, index order [x][y])
//   utils/visualization.py:130-134     (heat-map, bins=resolution, np.linspace edges)
//   app_simplified.py:205-209          (heat-map, bins=100)
//
// numpy semantics restated: per axis b = searchsorted(edges, x, 'right');
// x == edges[-1] goes to the last bin; everything outside (and NaN) is dropped.  Edges and compares
// are fp64; the fp32 inputs of the F32X4 layout are widened exactly.  The bin is found by a uniform
// guess followed by a correction walk against the *actual* edge values (Appendix A.3), with a
// binary search as the fallback for non-uniform edges, so any monotone edge array works.
//
// Two accumulation strategies, both order independent (integer adds) and therefore bit-exact:
//   GLOBAL  one RED.ADD.S32 per point into the grid, which is tiny (<= a few MB) and L2 resident.
//   SHARED  CTA-private grid in shared memory (<= 227 KB), atomics aggregated across the warp with
//           match.any, non-zero bins flushed with RED at the end.  Wins when points per CTA are
//           many times the bin count (scan-ordered frames, 50 M-point scans).
// HBM-bound: 16 B/point (F32X4) read once, nothing else.
#include "common.cuh"

namespace lidar {

constexpr int kHistThreads = 256;

struct EdgeView {
    const double* e;  // n+1 edges (shared or global memory)
    int n;            // bins
    double a;         // e[0]
    double inv_d;     // n / (e[n] - e[0])
};

__device__ __forceinline__ int find_bin(double x, const EdgeView& E) {
    const double lo = E.e[0], hi = E.e[E.n];
    if (!(x >= lo) || !(x <= hi)) return -1;  // also drops NaN
    if (x == hi) return E.n - 1;
    int k = (int)floor(__dmul_rn(__dsub_rn(x, E.a), E.inv_d));
    k = k < 0 ? 0 : (k > E.n - 1 ? E.n - 1 : k);
    int steps = 0;
    while (x < E.e[k]) {  // k > 0 guaranteed because x >= e[0]
        --k;
        if (++steps > 4) goto bsearch;
    }
    while (x >= E.e[k + 1]) {  // k < n-1 guaranteed because x < e[n]
        ++k;
        if (++steps > 8) goto bsearch;
    }
    return k;
bsearch : {
    // upper_bound(e, x) - 1
    int l = 0, r = E.n;  // invariant: e[l] <= x < e[r]
    while (r - l > 1) {
        int m = (l + r) >> 1;
        if (x >= E.e[m]) l = m; else r = m;
    }
    return l;
}
}

template <class Loader, bool kShared>
__global__ void __launch_bounds__(kHistThreads)
hist2d_kernel(Loader L, int64_t n, const double* __restrict__ g_ex, int nx,
              const double* __restrict__ g_ey, int ny, int32_t* __restrict__ counts,
              int edges_in_smem) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* s_edges = reinterpret_cast<double*>(smem_raw);
    EdgeView EX, EY;
    size_t edge_bytes = 0;
    if (edges_in_smem) {
        for (int i = threadIdx.x; i <= nx; i += blockDim.x) s_edges[i] = g_ex[i];
        for (int i = threadIdx.x; i <= ny; i += blockDim.x) s_edges[nx + 1 + i] = g_ey[i];
        EX.e = s_edges;
        EY.e = s_edges + nx + 1;
        edge_bytes = (size_t)(nx + ny + 2) * sizeof(double);
    } else {
        EX.e = g_ex;
        EY.e = g_ey;
    }
    int* s_hist = reinterpret_cast<int*>(smem_raw + ((edge_bytes + 15) & ~size_t(15)));
    const int nbins = nx * ny;
    if (kShared) {
        for (int i = threadIdx.x; i < nbins; i += blockDim.x) s_hist[i] = 0;
    }
    __syncthreads();
    EX.n = nx; EX.a = EX.e[0]; EX.inv_d = (double)nx / (EX.e[nx] - EX.e[0]);
    EY.n = ny; EY.a = EY.e[0]; EY.inv_d = (double)ny / (EY.e[ny] - EY.e[0]);

    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t start = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    // round the trip count up so the whole warp stays converged for match.any
    const int64_t n_round = ((n + 31) / 32) * 32;
    int64_t i = start;
    if (!kShared) {
        // GLOBAL: four independent loads and four fire-and-forget REDs in flight per thread.  (Measured on the
        // 50 M-point scan: the kernel runs at the L2 reduction rate, ~105 G RED/s; spreading the REDs over 8
        // private replicas of the grid changes it by 2 %, so there is no replication here.)
        for (; i + 3 * stride < n; i += 4 * stride) {
            Pt p[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) p[u] = L.load(i + u * stride);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int bx = find_bin(p[u].x, EX);
                const int by = find_bin(p[u].y, EY);
                if (bx >= 0 && by >= 0) atomicAdd(&counts[bx * ny + by], 1);   // result unused -> RED.E.ADD
            }
        }
    }
    for (; i < n_round; i += stride) {
        int bin = -1;
        if (i < n) {
            Pt p = L.load(i);
            const int bx = find_bin(p.x, EX);
            const int by = find_bin(p.y, EY);
            if (bx >= 0 && by >= 0) bin = bx * ny + by;
        }
        if (kShared) {
            const unsigned act = __activemask();
            const unsigned peers = __match_any_sync(act, bin);
            if (bin >= 0 && (peers & lanemask_lt()) == 0) atomicAdd(&s_hist[bin], __popc(peers));
        } else {
            if (bin >= 0) atomicAdd(&counts[bin], 1);  // result unused -> RED.E.ADD
        }
    }
    if (kShared) {
        __syncthreads();
        for (int i = threadIdx.x; i < nbins; i += blockDim.x) {
            const int c = s_hist[i];
            if (c) atomicAdd(&counts[i], c);
        }
    }
}

template <class Loader>
static int launch_hist2d(Loader L, int64_t n, const double* d_ex, int nx, const double* d_ey, int ny,
                         int32_t* d_counts, int mode, cudaStream_t st) {
    if (n == 0) return LIDAR_OK;
    const size_t optin = smem_optin();
    const size_t edge_bytes = (size_t)(nx + ny + 2) * sizeof(double);
    const int edges_in_smem = edge_bytes <= 32 * 1024;
    const size_t edge_smem = edges_in_smem ? ((edge_bytes + 15) & ~size_t(15)) : 0;
    const size_t hist_bytes = (size_t)nx * ny * sizeof(int);
    const bool shared_fits = edge_smem + hist_bytes <= optin;
    int grid_cap = sm_count() * 8;
    int64_t want = (n + kHistThreads - 1) / kHistThreads;
    bool use_shared;
    if (mode == LIDAR_HIST_SHARED) {
        LIDAR_REQUIRE(shared_fits, LIDAR_ERR_INVALID,
                      "lidar_hist2d: %dx%d grid (%zu B) does not fit %zu B of shared memory", nx, ny,
                      hist_bytes, optin);
        use_shared = true;
    } else if (mode == LIDAR_HIST_GLOBAL) {
        use_shared = false;
    } else {
        // the private grid pays off once every CTA sees several points per bin it has to flush
        const int ctas = sm_count();
        use_shared = shared_fits && (n / ctas) >= 4 * (int64_t)nx * ny;
    }
    if (use_shared) {
        // one CTA per SM when the grid eats most of shared memory, more when it is small
        int per_sm = (int)(optin / (edge_smem + hist_bytes + 1024));
        if (per_sm < 1) per_sm = 1;
        if (per_sm > 4) per_sm = 4;
        grid_cap = sm_count() * per_sm;
        const int grid = (int)(want < grid_cap ? want : grid_cap);
        auto kern = hist2d_kernel<Loader, true>;
        LIDAR_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)(edge_smem + hist_bytes)));
        kern<<<grid, kHistThreads, edge_smem + hist_bytes, st>>>(L, n, d_ex, nx, d_ey, ny, d_counts,
                                                                 edges_in_smem);
    } else {
        const int grid = (int)(want < grid_cap ? want : grid_cap);
        hist2d_kernel<Loader, false><<<grid, kHistThreads, edge_smem, st>>>(L, n, d_ex, nx, d_ey, ny,
                                                                            d_counts, edges_in_smem);
    }
    LIDAR_CHECK_LAUNCH();
    return LIDAR_OK;
}

}  // namespace lidar

using namespace lidar;

extern "C" {

int lidar_hist2d_f64(const double* d_u, int64_t su, const double* d_v, int64_t sv, int64_t n,
                     const double* d_ex, int nx, const double* d_ey, int ny, int32_t* d_counts,
                     int mode, void* stream) {
    LIDAR_REQUIRE(n >= 0 && nx > 0 && ny > 0, LIDAR_ERR_INVALID, "lidar_hist2d_f64: bad sizes n=%lld nx=%d ny=%d",
                  (long long)n, nx, ny);
    LIDAR_REQUIRE(d_ex && d_ey && d_counts, LIDAR_ERR_INVALID, "lidar_hist2d_f64: NULL argument");
    LIDAR_REQUIRE(n == 0 || (d_u && d_v), LIDAR_ERR_INVALID, "lidar_hist2d_f64: NULL coordinates");
    LIDAR_REQUIRE((int64_t)nx * ny < (1ll << 31), LIDAR_ERR_INVALID, "lidar_hist2d_f64: grid too large");
    return launch_hist2d(LoadF64uv{d_u, d_v, su, sv}, n, d_ex, nx, d_ey, ny, d_counts, mode,
                         as_stream(stream));
}

int lidar_hist2d_points(const void* d_points, int fmt, int64_t n, const double* d_ex, int nx,
                        const double* d_ey, int ny, int32_t* d_counts, int mode, void* stream) {
    LIDAR_REQUIRE(n >= 0 && nx > 0 && ny > 0, LIDAR_ERR_INVALID, "lidar_hist2d_points: bad sizes");
    LIDAR_REQUIRE(d_ex && d_ey && d_counts, LIDAR_ERR_INVALID, "lidar_hist2d_points: NULL argument");
    LIDAR_REQUIRE(n == 0 || d_points, LIDAR_ERR_INVALID, "lidar_hist2d_points: NULL points");
    LIDAR_REQUIRE((int64_t)nx * ny < (1ll << 31), LIDAR_ERR_INVALID, "lidar_hist2d_points: grid too large");
    if (fmt == LIDAR_FMT_F32X4)
        return launch_hist2d(LoadF32x4{static_cast<const float4*>(d_points)}, n, d_ex, nx, d_ey, ny,
                             d_counts, mode, as_stream(stream));
    if (fmt == LIDAR_FMT_F64X3)
        return launch_hist2d(LoadF64x3{static_cast<const double*>(d_points)}, n, d_ex, nx, d_ey, ny,
                             d_counts, mode, as_stream(stream));
    LIDAR_REQUIRE(false, LIDAR_ERR_INVALID, "lidar_hist2d_points: unknown point format %d", fmt);
}

}  // extern "C"
