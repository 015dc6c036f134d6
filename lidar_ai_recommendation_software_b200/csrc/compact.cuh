// Single-pass order-preserving stream compaction skeleton shared by roi_crop (compact.cu), the
// 3-sigma inlier filter and the ground split (preprocess.cu).
//
// A tile is kCmpTile consecutive points; thread t of the CTA owns points
// tile*kCmpTile + j*kCmpThreads + t (striped: coalesced loads).  The output slot of a kept point =
// exclusive prefix of the tile (chained scan, common.cuh) + #kept before it inside the tile, counted
// in index order from warp ballots.
#pragma once
#include "common.cuh"

namespace lidar {

constexpr int kCmpThreads = 256;
constexpr int kCmpRows = 8;  // points per thread
constexpr int kCmpTile = kCmpThreads * kCmpRows;

struct CompactCtrl {
    unsigned int ticket;
    unsigned int pad[3];
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


struct CompactLayout {
    size_t off_ctrl, off_desc, total;
    int64_t tiles;
};
static inline CompactLayout compact_layout(int64_t n) {
    CompactLayout L;
    L.tiles = (n + kCmpTile - 1) / kCmpTile;
    if (L.tiles < 1) L.tiles = 1;
    L.off_ctrl = 0;
    L.off_desc = ws_align(sizeof(CompactCtrl));
    L.total = ws_align(L.off_desc + sizeof(unsigned long long) * L.tiles);
    return L;
}


}  // namespace lidar
