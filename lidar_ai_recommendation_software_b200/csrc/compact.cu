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
#include "common.cuh"

namespace lidar {

constexpr int kCmpThreads = 256;
constexpr int kCmpRows = 8;  // points per thread
constexpr int kCmpTile = kCmpThreads * kCmpRows;

struct CompactCtrl {
    unsigned int ticket;
    unsigned int pad[3];
};

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

// Shared skeleton: `keep[j]` flags for this thread's kCmpRows points of `tile` -> output slots.
// Returns in slot[j] the global output index (or -1).  All threads of the CTA must call it.
__device__ __forceinline__ void compact_slots(const bool (&keep)[kCmpRows], int tile,
                                              unsigned long long* tile_desc, long long (&slot)[kCmpRows],
                                              unsigned long long* total_out, bool is_last_tile) {
    __shared__ unsigned s_cnt[kCmpRows][kCmpThreads / 32];
    __shared__ unsigned long long s_base;
    const unsigned lane = lane_id();
    const int warp = threadIdx.x >> 5;
    unsigned ballots[kCmpRows];
#pragma unroll
    for (int j = 0; j < kCmpRows; ++j) {
        ballots[j] = __ballot_sync(0xffffffffu, keep[j]);
        if (lane == 0) s_cnt[j][warp] = __popc(ballots[j]);
    }
    __syncthreads();
    // exclusive prefix over (row-major j, warp) = index order inside the tile
    unsigned total = 0;
    unsigned my_off[kCmpRows];
#pragma unroll
    for (int j = 0; j < kCmpRows; ++j) {
#pragma unroll
        for (int w = 0; w < kCmpThreads / 32; ++w) {
            if (w == warp) my_off[j] = total;
            total += s_cnt[j][w];
        }
    }
    if (warp == 0) {
        const unsigned long long ex = scan_lookback_warp(tile_desc, tile, (unsigned long long)total);
        if (lane == 0) {
            s_base = ex;
            if (is_last_tile && total_out) *total_out = ex + total;
        }
    }
    __syncthreads();
    const unsigned long long base = s_base;
#pragma unroll
    for (int j = 0; j < kCmpRows; ++j) {
        slot[j] = keep[j] ? (long long)(base + my_off[j] + __popc(ballots[j] & lanemask_lt())) : -1ll;
    }
    __syncthreads();  // s_cnt / s_base are reused by the next tile
}

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

struct CompactLayout {
    size_t off_ctrl, off_desc, total;
    int64_t tiles;
};
static CompactLayout compact_layout(int64_t n) {
    CompactLayout L;
    L.tiles = (n + kCmpTile - 1) / kCmpTile;
    if (L.tiles < 1) L.tiles = 1;
    L.off_ctrl = 0;
    L.off_desc = ws_align(sizeof(CompactCtrl));
    L.total = ws_align(L.off_desc + sizeof(unsigned long long) * L.tiles);
    return L;
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
