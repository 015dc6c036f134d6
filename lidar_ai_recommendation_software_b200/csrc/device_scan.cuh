// Exclusive prefix sum of a uint32 array in one pass (chained scan with look-back, common.cuh).
// Used for cell-list offsets and cluster numbering (dbscan.cu) and segment offsets (cluster.cu).
#pragma once
#include "common.cuh"

namespace lidar {

constexpr int kDsThreads = 256;
constexpr int kDsItems = 16;
constexpr int kDsTile = kDsThreads * kDsItems;

struct ScanCtrl {
    unsigned int ticket;
    unsigned int pad[3];
};

static inline int64_t scan_tiles(int64_t n) { return n <= 0 ? 1 : (n + kDsTile - 1) / kDsTile; }
static inline size_t scan_workspace_bytes(int64_t n) {
    return ws_align(sizeof(ScanCtrl)) + ws_align(sizeof(unsigned long long) * scan_tiles(n));
}

// out[i] = sum(in[0..i)), out[n] = total (out has n+1 entries); *total64 (optional) = total
template <class T>
__global__ void __launch_bounds__(kDsThreads)
exclusive_scan_kernel(const T* __restrict__ in, unsigned* __restrict__ out, int64_t n,
                      unsigned long long* __restrict__ total64, unsigned long long* tile_desc, ScanCtrl* ctrl,
                      int n_tiles) {
    __shared__ int s_tile;
    __shared__ unsigned s_warp[kDsThreads / 32];
    __shared__ unsigned long long s_excl;
    const unsigned lane = lane_id();
    const int warp = threadIdx.x >> 5;
    while (true) {
        if (threadIdx.x == 0) s_tile = (int)atomicAdd(&ctrl->ticket, 1u);
        __syncthreads();
        const int tile = s_tile;
        if (tile >= n_tiles) break;
        // blocked arrangement: thread t owns items [t*8, t*8+8) of the tile
        const int64_t base = (int64_t)tile * kDsTile + (int64_t)threadIdx.x * kDsItems;
        unsigned v[kDsItems];
        unsigned sum = 0;
        // whole tiles of 4-byte items move as 16-byte vectors (four loads in flight per thread; the scalar form with
        // its per-item bound checks ran at a fifth of the memory rate: 24 us for the 3.6 M-cell directory of a ring frame)
        const bool vec = sizeof(T) == 4 && (int64_t)tile * kDsTile + kDsTile <= n &&
                         ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0;
        if (vec) {
            const uint4* src = reinterpret_cast<const uint4*>(in + base);
            uint4 q[kDsItems / 4];
#pragma unroll
            for (int k = 0; k < kDsItems / 4; ++k) q[k] = __ldg(src + k);
#pragma unroll
            for (int k = 0; k < kDsItems / 4; ++k) {
                v[4 * k] = q[k].x; v[4 * k + 1] = q[k].y; v[4 * k + 2] = q[k].z; v[4 * k + 3] = q[k].w;
            }
#pragma unroll
            for (int k = 0; k < kDsItems; ++k) sum += v[k];
        } else {
#pragma unroll
            for (int k = 0; k < kDsItems; ++k) {
                v[k] = (base + k < n) ? (unsigned)in[base + k] : 0u;
                sum += v[k];
            }
        }
        unsigned inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= (unsigned)o) inc += t;
        }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        unsigned warp_off = 0, total = 0;
#pragma unroll
        for (int w = 0; w < kDsThreads / 32; ++w) {
            const unsigned s = s_warp[w];
            if (w < warp) warp_off += s;
            total += s;
        }
        if (warp == 0) {
            const unsigned long long ex = scan_lookback_warp(tile_desc, tile, (unsigned long long)total);
            if (lane == 0) {
                s_excl = ex;
                if (tile == n_tiles - 1) {
                    out[n] = (unsigned)(ex + total);
                    if (total64) *total64 = ex + total;
                }
            }
        }
        __syncthreads();
        unsigned run = (unsigned)s_excl + warp_off + (inc - sum);
        if (vec) {
            uint4* dst = reinterpret_cast<uint4*>(out + base);
#pragma unroll
            for (int k = 0; k < kDsItems / 4; ++k) {
                uint4 o;
                o.x = run; run += v[4 * k];
                o.y = run; run += v[4 * k + 1];
                o.z = run; run += v[4 * k + 2];
                o.w = run; run += v[4 * k + 3];
                dst[k] = o;
            }
        } else {
#pragma unroll
            for (int k = 0; k < kDsItems; ++k) {
                if (base + k < n) out[base + k] = run;
                run += v[k];
            }
        }
        __syncthreads();
    }
}

// d_ws must hold scan_workspace_bytes(n) bytes.  Enqueues a memset of the control block + the scan.
template <class T>
static inline cudaError_t launch_exclusive_scan(const T* in, unsigned* out, int64_t n, unsigned long long* total64,
                                                void* d_ws, cudaStream_t st) {
    char* ws = static_cast<char*>(d_ws);
    const int64_t tiles = scan_tiles(n);
    cudaError_t e = cudaMemsetAsync(ws, 0, scan_workspace_bytes(n), st);
    if (e != cudaSuccess) return e;
    int grid = sm_count() * 4;
    if ((int64_t)grid > tiles) grid = (int)tiles;
    exclusive_scan_kernel<T><<<grid, kDsThreads, 0, st>>>(
        in, out, n, total64, reinterpret_cast<unsigned long long*>(ws + ws_align(sizeof(ScanCtrl))),
        reinterpret_cast<ScanCtrl*>(ws), (int)tiles);
    return cudaGetLastError();
}

}  // namespace lidar
