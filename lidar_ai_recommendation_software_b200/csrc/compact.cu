// a5 — ROI crop: predicate + order-preserving stream compaction in ONE pass over the points.
//
// NEW op (no reference call site; contract = SURVEY.md Appendix B.2): keep <=> lo <= p <= hi on
// x, y and z.  The same single-pass compaction skeleton serves the 3-sigma inlier filter of the
// preprocess stage (utils/data_processing.py:155-157, see preprocess.cu).
//
// Layout: a tile is kTileItems consecutive points; thread t of the CTA owns points
// tile*kTileItems + j*blockDim + t (striped, so loads are coalesced LDG.128s).  The output rank
// of a kept point = exclusive prefix of the tile (chained scan, common.cuh) + #kept before it in
// the tile, counted in index order from warp ballots.  Reads 16 B/point, writes 16 B per kept
// point + 1 B mask: HBM-bound.
#include "compact.cuh"

namespace lidar {

struct BoxPredF32 {
    float lo[3], hi[3];
    __device__ __forceinline__ bool operator()(const float4& v) const {
        return v.x >= lo[0] && v.x <= hi[0] && v.y >= lo[1] && v.y <= hi[1] && v.z >= lo[2] && v.z <= hi[2];
    }
};
struct BoxPredF64 {
    double lo[3], hi[3];
    __device__ __forceinline__ bool operator()(double x, double y, double z) const {
        return x >= lo[0] && x <= hi[0] && y >= lo[1] && y <= hi[1] && z >= lo[2] && z <= hi[2];
    }
};

__global__ void __launch_bounds__(kCmpThreads)
roi_crop_f32x4_kernel(const float4* __restrict__ pts, int64_t n, BoxPredF32 pred, uint8_t* __restrict__ mask,
                      float4* __restrict__ out, int64_t* __restrict__ count, unsigned long long* tile_desc,
                      CompactCtrl* ctrl, int n_tiles) {
    __shared__ int s_tile;
    LoadF32x4 L{pts};
    while (true) {
        if (threadIdx.x == 0) s_tile = (int)atomicAdd(&ctrl->ticket, 1u);
        __syncthreads();
        const int tile = s_tile;
        if (tile >= n_tiles) break;
        const int64_t base = (int64_t)tile * kCmpTile;
        float4 v[kCmpRows];
        bool keep[kCmpRows];
#pragma unroll
        for (int j = 0; j < kCmpRows; ++j) {
            const int64_t i = base + (int64_t)j * kCmpThreads + threadIdx.x;
            keep[j] = false;
            if (i < n) {
                v[j] = L.raw(i);
                keep[j] = pred(v[j]);
                if (mask) mask[i] = keep[j] ? 1 : 0;
            }
        }
        long long slot[kCmpRows];
        compact_slots(keep, tile, tile_desc, slot, reinterpret_cast<unsigned long long*>(count), tile == n_tiles - 1);
#pragma unroll
        for (int j = 0; j < kCmpRows; ++j)
            if (slot[j] >= 0) out[slot[j]] = v[j];
    }
}

__global__ void __launch_bounds__(kCmpThreads)
roi_crop_f64x3_kernel(const double* __restrict__ pts, int64_t n, BoxPredF64 pred, uint8_t* __restrict__ mask,
                      double* __restrict__ out, int64_t* __restrict__ count, unsigned long long* tile_desc,
                      CompactCtrl* ctrl, int n_tiles) {
    __shared__ int s_tile;
    while (true) {
        if (threadIdx.x == 0) s_tile = (int)atomicAdd(&ctrl->ticket, 1u);
        __syncthreads();
        const int tile = s_tile;
        if (tile >= n_tiles) break;
        const int64_t base = (int64_t)tile * kCmpTile;
        double x[kCmpRows], y[kCmpRows], z[kCmpRows];
        bool keep[kCmpRows];
#pragma unroll
        for (int j = 0; j < kCmpRows; ++j) {
            const int64_t i = base + (int64_t)j * kCmpThreads + threadIdx.x;
            keep[j] = false;
            if (i < n) {
                x[j] = __ldg(pts + 3 * i);
                y[j] = __ldg(pts + 3 * i + 1);
                z[j] = __ldg(pts + 3 * i + 2);
                keep[j] = pred(x[j], y[j], z[j]);
                if (mask) mask[i] = keep[j] ? 1 : 0;
            }
        }
        long long slot[kCmpRows];
        compact_slots(keep, tile, tile_desc, slot, reinterpret_cast<unsigned long long*>(count), tile == n_tiles - 1);
#pragma unroll
        for (int j = 0; j < kCmpRows; ++j)
            if (slot[j] >= 0) {
                double* o = out + 3 * slot[j];
                o[0] = x[j]; o[1] = y[j]; o[2] = z[j];
            }
    }
}

}  // namespace lidar

using namespace lidar;

extern "C" {

size_t lidar_compact_workspace_bytes(int64_t n) { return compact_layout(n < 0 ? 0 : n).total; }

int lidar_roi_crop(const void* d_points, int fmt, int64_t n, const double* h_lo3, const double* h_hi3,
                   uint8_t* d_mask, void* d_out, int64_t* d_count, void* d_ws, size_t ws_bytes,
                   void* stream) {
    LIDAR_REQUIRE(n >= 0, LIDAR_ERR_INVALID, "lidar_roi_crop: n < 0");
    LIDAR_REQUIRE(h_lo3 && h_hi3 && d_count, LIDAR_ERR_INVALID, "lidar_roi_crop: NULL argument");
    LIDAR_REQUIRE(n == 0 || (d_points && d_out), LIDAR_ERR_INVALID, "lidar_roi_crop: NULL points");
    LIDAR_REQUIRE(n < (int64_t)kCmpTile * 0x7fffffff, LIDAR_ERR_INVALID, "lidar_roi_crop: n too large");
    const CompactLayout L = compact_layout(n);
    LIDAR_REQUIRE(d_ws && ws_bytes >= L.total, LIDAR_ERR_WORKSPACE,
                  "lidar_roi_crop: workspace too small (%zu < %zu)", ws_bytes, L.total);
    cudaStream_t st = as_stream(stream);
    LIDAR_CUDA_TRY(cudaMemsetAsync(d_ws, 0, L.total, st));
    LIDAR_CUDA_TRY(cudaMemsetAsync(d_count, 0, sizeof(int64_t), st));
    if (n == 0) return LIDAR_OK;
    char* ws = static_cast<char*>(d_ws);
    CompactCtrl* ctrl = reinterpret_cast<CompactCtrl*>(ws + L.off_ctrl);
    unsigned long long* desc = reinterpret_cast<unsigned long long*>(ws + L.off_desc);
    int grid = sm_count() * 4;
    if ((int64_t)grid > L.tiles) grid = (int)L.tiles;
    if (fmt == LIDAR_FMT_F32X4) {
        BoxPredF32 p;
        for (int c = 0; c < 3; ++c) {
            // fp32 compares (B.2): the bounds are rounded to fp32 toward the inside of the box? No —
            // the contract compares fp32(p) with fp32(bound); the oracle does exactly the same cast.
            p.lo[c] = (float)h_lo3[c];
            p.hi[c] = (float)h_hi3[c];
        }
        roi_crop_f32x4_kernel<<<grid, kCmpThreads, 0, st>>>(static_cast<const float4*>(d_points), n, p, d_mask,
                                                            static_cast<float4*>(d_out), d_count, desc, ctrl,
                                                            (int)L.tiles);
    } else if (fmt == LIDAR_FMT_F64X3) {
        BoxPredF64 p;
        for (int c = 0; c < 3; ++c) {
            p.lo[c] = h_lo3[c];
            p.hi[c] = h_hi3[c];
        }
        roi_crop_f64x3_kernel<<<grid, kCmpThreads, 0, st>>>(static_cast<const double*>(d_points), n, p, d_mask,
                                                            static_cast<double*>(d_out), d_count, desc, ctrl,
                                                            (int)L.tiles);
    } else {
        LIDAR_REQUIRE(false, LIDAR_ERR_INVALID, "lidar_roi_crop: unknown point format %d", fmt);
    }
    LIDAR_CHECK_LAUNCH();
    return LIDAR_OK;
}

}  // extern "C"
