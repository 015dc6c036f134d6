// K5, general path — voxel downsample by a deterministic radix SORT of 64-bit voxel keys + segmented reduction
// (the north star's formulation).
//
// The occupancy-bitmap back ends (voxel.cu) hold one bit per cell of the key space and stop at 2^31 cells: one outlier
// return, or a 1 km x 1 km x 30 m venue at 0.05 m (2.4e11 cells), is beyond them.  This path has no such limit:
//   bbox (of the points inside the optional ROI box) -> origin, dims, key width     [device-side descriptor]
//   key64 = (ix*Dy + iy)*Dz + iz per point, same exact index arithmetic as voxel.cu  (cropped points: key = ~0)
//   LSD radix sort of (key, index) pairs, 8 bits per pass, ONLY the passes the key width needs (read on the device:
//     the launches of unneeded passes return at once), stable: members stay in ascending original index
//   head flags + chained scan -> rank of every voxel, unique keys, segment starts, inverse
//   one thread per voxel: exact fixed-point sums of its members (the arithmetic of the bitmap path: (p - ref) * 2^k is
//     an exact integer), centroid = fp32(ref + sum / count), count
// Nothing visits the host; the descriptor comes back with the results.  Outputs equal the bitmap path's wherever both
// apply (same keys, ranks, counts; centroids from the same exact sums), and the int64-key oracle everywhere.
#include "common.cuh"

namespace lidar {

constexpr int kSortThreads = 256;
constexpr int kSortItems = 4;                           // keys per thread per tile
constexpr int kSortTile = kSortThreads * kSortItems;    // 1024 keys per block
constexpr unsigned long long kCropKey = ~0ull;

struct SortWs {
    double partial[1024][8];
    unsigned ticket;
    unsigned pad[3];
    unsigned long long scan_state[1];   // followed by the chained-scan descriptors (one per tile)
};

struct SortLayout {
    size_t off_ws, off_desc_tiles, off_keys_a, off_keys_b, off_idx_a, off_idx_b, off_hist, off_rank, off_start, total;
    int64_t tiles;
};

static SortLayout sort_layout(int64_t n) {
    SortLayout L;
    L.tiles = (n + kSortTile - 1) / kSortTile;
    if (L.tiles < 1) L.tiles = 1;
    size_t o = 0;
    auto take = [&](size_t b) { size_t a = ws_align(o); o = a + b; return a; };
    L.off_ws = take(sizeof(SortWs));
    L.off_desc_tiles = take(sizeof(unsigned long long) * (size_t)L.tiles);
    L.off_keys_a = take(sizeof(unsigned long long) * (size_t)(n > 0 ? n : 1));
    L.off_keys_b = take(sizeof(unsigned long long) * (size_t)(n > 0 ? n : 1));
    L.off_idx_a = take(sizeof(unsigned) * (size_t)(n > 0 ? n : 1));
    L.off_idx_b = take(sizeof(unsigned) * (size_t)(n > 0 ? n : 1));
    L.off_hist = take(sizeof(unsigned) * 256 * (size_t)L.tiles);
    L.off_rank = take(sizeof(unsigned) * (size_t)(n > 0 ? n : 1));
    L.off_start = take(sizeof(unsigned) * (size_t)(n + 1));
    L.total = ws_align(o);
    return L;
}

struct SortParams {
    const float4* pts;
    int64_t n;
    double voxel;
    double origin[3];
    int has_origin;
    float lo[3], hi[3];
    int has_roi;
};

__device__ __forceinline__ bool in_roi(const float4& q, const SortParams& P) {
    return !P.has_roi || (q.x >= P.lo[0] && q.x <= P.hi[0] && q.y >= P.lo[1] && q.y <= P.hi[1] && q.z >= P.lo[2] && q.z <= P.hi[2]);
}

__device__ __forceinline__ long long sorted_floor_div(double d, double v, double rinv) {
    // floor(fl(d / v)): reciprocal guess, exact division when the guess is within 4 ulp of an integer (voxel.cu)
    const double qh = __dmul_rn(d, rinv);
    const double fl = floor(qh);
    const double fr = __dsub_rn(qh, fl);
    const double tol = __dadd_rn(__dmul_rn(fabs(qh), 0x1p-50), 1e-300);
    if (fr > tol && __dsub_rn(1.0, fr) > tol) return (long long)fl;
    return (long long)floor(__ddiv_rn(d, v));
}

// ---- bbox of the kept points + descriptor --------------------------------------------------------------------
__global__ void __launch_bounds__(kSortThreads)
k_sorted_bbox(SortParams P, SortWs* ws, lidar_sorted_desc* D, int fix_bits_budget) {
    float mn[4] = {INFINITY, INFINITY, INFINITY, INFINITY}, mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    unsigned long long kept = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P.n; i += stride) {
        const float4 q = __ldg(P.pts + i);
        if (in_roi(q, P)) {
            mn[0] = fminf(mn[0], q.x); mx[0] = fmaxf(mx[0], q.x);
            mn[1] = fminf(mn[1], q.y); mx[1] = fmaxf(mx[1], q.y);
            mn[2] = fminf(mn[2], q.z); mx[2] = fmaxf(mx[2], q.z);
            mn[3] = fminf(mn[3], q.w); mx[3] = fmaxf(mx[3], q.w);
            ++kept;
        }
    }
    __shared__ double s_v[kSortThreads / 32][9];
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
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) kept += __shfl_xor_sync(0xffffffffu, kept, o);
    if (lane_id() == 0) {
#pragma unroll
        for (int c = 0; c < 4; ++c) { s_v[warp][c] = mn[c]; s_v[warp][4 + c] = mx[c]; }
        s_v[warp][8] = (double)kept;
    }
    __syncthreads();
    if (threadIdx.x < 9) {
        double v = s_v[0][threadIdx.x];
        for (int w = 1; w < kSortThreads / 32; ++w)
            v = threadIdx.x == 8 ? v + s_v[w][8] : threadIdx.x >= 4 ? fmax(v, s_v[w][threadIdx.x]) : fmin(v, s_v[w][threadIdx.x]);
        if (threadIdx.x < 8) ws->partial[blockIdx.x][threadIdx.x] = v;
        else atomicAdd(reinterpret_cast<unsigned long long*>(&D->n_kept), (unsigned long long)v);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(&ws->ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (threadIdx.x == 0) {
        double bb[8];
        for (int c = 0; c < 8; ++c) bb[c] = c < 4 ? INFINITY : -INFINITY;
        for (unsigned b = 0; b < gridDim.x; ++b)
            for (int c = 0; c < 8; ++c) {
                const double q = __ldcg(&ws->partial[b][c]);
                bb[c] = c < 4 ? fmin(bb[c], q) : fmax(bb[c], q);
            }
        const long long n_kept = (long long)__ldcg(reinterpret_cast<unsigned long long*>(&D->n_kept));
        int status = 0;
        long long ks = 1;
        for (int c = 0; c < 3; ++c) {
            const double o = P.has_origin ? P.origin[c] : bb[c];
            D->origin[c] = o;
            D->bbox_min[c] = bb[c];
            D->bbox_max[c] = bb[4 + c];
            long long d = 1;
            if (n_kept > 0) {
                if (bb[c] < o) status = LIDAR_ERR_INVALID;                     // a point below the origin would index < 0
                const double span = floor(__ddiv_rn(__dsub_rn(bb[4 + c], o), P.voxel));
                d = (span >= 0.0 && span < 9.0e15) ? (long long)span + 1 : 0;
                if (d <= 0) { d = 1; status = status ? status : LIDAR_ERR_CAPACITY; }
            }
            D->dims[c] = d;
            // key space must stay below 2^63 (the crop sentinel is 2^64 - 1)
            if (ks > 0 && d > (long long)(0x7fffffffffffffffll / ks)) { status = status ? status : LIDAR_ERR_CAPACITY; ks = 0; }
            else ks *= d;
        }
        D->key_space = ks;
        // digits looked at: enough that 2^(8 passes) - 1 -- what the crop sentinel ~0 looks like to those passes -- is
        // strictly above the largest real key, so cropped points sort behind every kept one
        int bits = 0;
        while (bits < 63 && (1ll << bits) <= ks) ++bits;
        D->passes = (n_kept > 0 && status == 0) ? (bits + 7) / 8 : 0;
        D->voxel = P.voxel;
        D->n_points = P.n;
        D->n_voxels = 0;
        int e;
        frexp(P.voxel * 2.0 + 0x1p-24, &e);
        D->fix_scale_xyz = ldexp(1.0, fix_bits_budget - e);
        double wmax = fmax(fabs(bb[3]), fabs(bb[7]));
        if (!(wmax > 0.0) || !isfinite(wmax)) wmax = 1.0;
        frexp(wmax, &e);
        D->fix_scale_w = ldexp(1.0, fix_bits_budget - e - 1);
        D->status = status;
        ws->ticket = 0u;
    }
}

// ---- keys ----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSortThreads)
k_sorted_keys(SortParams P, const lidar_sorted_desc* __restrict__ Dg, unsigned long long* __restrict__ keys,
              unsigned* __restrict__ idx, long long* __restrict__ voxel_key) {
    __shared__ lidar_sorted_desc D;
    if (threadIdx.x == 0) D = *Dg;
    __syncthreads();
    const double rinv = __ddiv_rn(1.0, D.voxel);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P.n; i += stride) {
        const float4 q = __ldg(P.pts + i);
        unsigned long long k = kCropKey;
        if (D.status == 0 && in_roi(q, P)) {
            const long long ix = sorted_floor_div(__dsub_rn((double)q.x, D.origin[0]), D.voxel, rinv);
            const long long iy = sorted_floor_div(__dsub_rn((double)q.y, D.origin[1]), D.voxel, rinv);
            const long long iz = sorted_floor_div(__dsub_rn((double)q.z, D.origin[2]), D.voxel, rinv);
            k = (unsigned long long)((ix * D.dims[1] + iy) * D.dims[2] + iz);
        }
        keys[i] = k;
        idx[i] = (unsigned)i;
        voxel_key[i] = k == kCropKey ? -1ll : (long long)k;
    }
}

// ---- LSD radix sort, one 8-bit digit per pass -----------------------------------------------------------------
__device__ __forceinline__ void sort_buffers(int pass, unsigned long long* ka, unsigned long long* kb, unsigned* ia, unsigned* ib,
                                             unsigned long long*& ksrc, unsigned long long*& kdst, unsigned*& isrc, unsigned*& idst) {
    if (pass & 1) { ksrc = kb; kdst = ka; isrc = ib; idst = ia; } else { ksrc = ka; kdst = kb; isrc = ia; idst = ib; }
}

__global__ void __launch_bounds__(kSortThreads)
k_sort_hist(int pass, int64_t n, const lidar_sorted_desc* __restrict__ Dg, unsigned long long* ka, unsigned long long* kb,
            unsigned* ia, unsigned* ib, unsigned* __restrict__ hist, int tiles) {
    if (pass >= Dg->passes) return;
    unsigned long long *ksrc, *kdst;
    unsigned *isrc, *idst;
    sort_buffers(pass, ka, kb, ia, ib, ksrc, kdst, isrc, idst);
    __shared__ unsigned s_h[256];
    s_h[threadIdx.x] = 0u;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * kSortTile;
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const int64_t i = base + r * kSortThreads + threadIdx.x;
        if (i < n) atomicAdd(&s_h[(unsigned)(ksrc[i] >> (8 * pass)) & 255u], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * tiles + blockIdx.x] = s_h[threadIdx.x];     // digit-major: a scan over it orders digits, then tiles
}

// exclusive scan of hist[256 * tiles] by ONE block (the fallback path is not the throughput path; 1 M points = 1 024 tiles)
__global__ void __launch_bounds__(1024)
k_sort_scan(int pass, const lidar_sorted_desc* __restrict__ Dg, unsigned* __restrict__ hist, int64_t count) {
    if (pass >= Dg->passes) return;
    __shared__ unsigned s_w[32];
    __shared__ unsigned s_carry;
    if (threadIdx.x == 0) s_carry = 0u;
    __syncthreads();
    const int64_t per = (count + 1023) / 1024;
    const int64_t i0 = threadIdx.x * per < count ? threadIdx.x * per : count;
    const int64_t i1 = i0 + per < count ? i0 + per : count;
    unsigned mine = 0;
    for (int64_t i = i0; i < i1; ++i) mine += hist[i];
    unsigned inc = mine;
    const unsigned lane = lane_id();
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (unsigned)o) inc += t;
    }
    if (lane == 31) s_w[threadIdx.x >> 5] = inc;
    __syncthreads();
    unsigned woff = 0;
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) woff += s_w[w];
    unsigned run = woff + inc - mine;
    for (int64_t i = i0; i < i1; ++i) {
        const unsigned v = hist[i];
        hist[i] = run;
        run += v;
    }
}

__global__ void __launch_bounds__(kSortThreads)
k_sort_scatter(int pass, int64_t n, const lidar_sorted_desc* __restrict__ Dg, unsigned long long* ka, unsigned long long* kb,
               unsigned* ia, unsigned* ib, const unsigned* __restrict__ hist, int tiles) {
    if (pass >= Dg->passes) return;
    unsigned long long *ksrc, *kdst;
    unsigned *isrc, *idst;
    sort_buffers(pass, ka, kb, ia, ib, ksrc, kdst, isrc, idst);
    __shared__ unsigned s_base[256];                       // next free slot of every digit for this tile
    __shared__ unsigned s_cnt[kSortThreads / 32][256];     // per warp, per digit: members in the current round
    s_base[threadIdx.x] = hist[(size_t)threadIdx.x * tiles + blockIdx.x];
    const int warp = threadIdx.x >> 5;
    const int64_t base = (int64_t)blockIdx.x * kSortTile;
    // rounds of 256 consecutive keys: within a round warps are ordered by warp id, lanes by lane id = index order
    for (int r = 0; r < kSortItems; ++r) {
        for (int k = threadIdx.x; k < (kSortThreads / 32) * 256; k += kSortThreads) (&s_cnt[0][0])[k] = 0u;
        __syncthreads();
        const int64_t i = base + r * kSortThreads + threadIdx.x;
        const bool live = i < n;
        unsigned long long key = 0ull;
        unsigned id = 0u, digit = 0u, rank_in_warp = 0u;
        if (live) { key = ksrc[i]; id = isrc[i]; digit = (unsigned)(key >> (8 * pass)) & 255u; }
        const unsigned act = __ballot_sync(0xffffffffu, live);
        if (live) {
            const unsigned peers = __match_any_sync(act, digit);
            rank_in_warp = __popc(peers & lanemask_lt());
            if (rank_in_warp == 0) s_cnt[warp][digit] = __popc(peers);
        }
        __syncthreads();
        if (live) {
            unsigned before = 0;
            for (int w = 0; w < warp; ++w) before += s_cnt[w][digit];
            const unsigned dest = s_base[digit] + before + rank_in_warp;
            kdst[dest] = key;
            idst[dest] = id;
        }
        __syncthreads();
        {
            unsigned tot = 0;
            for (int w = 0; w < kSortThreads / 32; ++w) tot += s_cnt[w][threadIdx.x];
            s_base[threadIdx.x] += tot;
        }
        __syncthreads();
    }
}

// ---- heads, ranks, unique keys, segment starts, inverse --------------------------------------------------------
__global__ void __launch_bounds__(kSortThreads)
k_sorted_rank(int64_t n, lidar_sorted_desc* __restrict__ Dg, const unsigned long long* ka, const unsigned long long* kb,
              const unsigned* ia, const unsigned* ib, unsigned long long* __restrict__ tile_desc, unsigned* __restrict__ ticket,
              unsigned* __restrict__ start, long long* __restrict__ unique_keys, int* __restrict__ inverse) {
    const int passes = Dg->passes;
    const unsigned long long* keys = (passes & 1) ? kb : ka;          // where the last pass left the pairs
    const unsigned* idx = (passes & 1) ? ib : ia;
    __shared__ unsigned s_w[kSortThreads / 32];
    __shared__ unsigned long long s_excl;
    __shared__ unsigned s_tile;
    // tiles are taken in ticket order, so the tile a block looks back at is always already running
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const unsigned tile = s_tile;
    const int64_t base = (int64_t)tile * kSortTile + (int64_t)threadIdx.x * kSortItems;
    unsigned long long k[kSortItems];
    unsigned head[kSortItems];
    unsigned mine = 0;
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const int64_t j = base + r;
        k[r] = j < n ? keys[j] : kCropKey;
        const unsigned long long prev = j == 0 ? kCropKey : (j - 1 < n ? keys[j - 1] : kCropKey);
        head[r] = (j < n && k[r] != kCropKey && (j == 0 || k[r] != prev)) ? 1u : 0u;
        mine += head[r];
    }
    unsigned inc = mine;
    const unsigned lane = lane_id();
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (unsigned)o) inc += t;
    }
    if (lane == 31) s_w[threadIdx.x >> 5] = inc;
    __syncthreads();
    unsigned woff = 0, total = 0;
    for (int w = 0; w < kSortThreads / 32; ++w) { if (w < (int)(threadIdx.x >> 5)) woff += s_w[w]; total += s_w[w]; }
    if (threadIdx.x < 32) {
        const unsigned long long e = scan_lookback_warp(tile_desc, (int)tile, (unsigned long long)total);
        if (threadIdx.x == 0) s_excl = e;
    }
    __syncthreads();
    unsigned run = (unsigned)s_excl + woff + inc - mine;       // voxels before this thread's first key
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const int64_t j = base + r;
        if (j < n && k[r] != kCropKey) {
            run += head[r];
            const unsigned rank = run - 1u;
            if (head[r]) { unique_keys[rank] = (long long)k[r]; start[rank] = (unsigned)j; }
            inverse[idx[j]] = (int)rank;
        } else if (j < n) {
            inverse[idx[j]] = -1;
        }
    }
    if (tile == gridDim.x - 1 && threadIdx.x == 0) {
        const long long V = (long long)(s_excl + total);
        Dg->n_voxels = V;
        start[V] = (unsigned)Dg->n_kept;                       // end of the last segment (kept keys sort before the crop sentinel)
    }
}

// ---- per-voxel exact sums -> centroid, count -------------------------------------------------------------------
__global__ void __launch_bounds__(kSortThreads)
k_sorted_reduce(const float4* __restrict__ pts, const lidar_sorted_desc* __restrict__ Dg, const unsigned* ia, const unsigned* ib,
                const unsigned* __restrict__ start, const long long* __restrict__ unique_keys, float4* __restrict__ centroids,
                int* __restrict__ counts) {
    __shared__ lidar_sorted_desc D;
    if (threadIdx.x == 0) D = *Dg;
    __syncthreads();
    if (D.status != 0) return;
    const unsigned* idx = (D.passes & 1) ? ib : ia;
    const double isx = 1.0 / D.fix_scale_xyz, isw = 1.0 / D.fix_scale_w;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < D.n_voxels; r += stride) {
        const long long key = unique_keys[r];
        const long long t = key / D.dims[2];
        const long long iz = key - t * D.dims[2];
        const long long ix = t / D.dims[1];
        const long long iy = t - ix * D.dims[1];
        // accumulation reference: the voxel corner snapped to 2^-24 m, so that (p - ref) * 2^k is an exact integer (voxel.cu)
        auto ref = [&](double o, long long i) {
            const double c = __dadd_rn(o, __dmul_rn((double)i, D.voxel));
            return __dmul_rn(nearbyint(__dmul_rn(c, 16777216.0)), 1.0 / 16777216.0);
        };
        const double cx = ref(D.origin[0], ix), cy = ref(D.origin[1], iy), cz = ref(D.origin[2], iz);
        long long sx = 0, sy = 0, sz = 0, sw = 0;
        const unsigned j0 = start[r], j1 = start[r + 1];
        for (unsigned j = j0; j < j1; ++j) {
            const float4 q = __ldg(pts + idx[j]);
            sx += __double2ll_rn(__dmul_rn(__dsub_rn((double)q.x, cx), D.fix_scale_xyz));
            sy += __double2ll_rn(__dmul_rn(__dsub_rn((double)q.y, cy), D.fix_scale_xyz));
            sz += __double2ll_rn(__dmul_rn(__dsub_rn((double)q.z, cz), D.fix_scale_xyz));
            sw += __double2ll_rn(__dmul_rn((double)q.w, D.fix_scale_w));
        }
        const double dc = (double)(j1 - j0);
        centroids[r] = make_float4((float)__dadd_rn(cx, __ddiv_rn(__dmul_rn((double)sx, isx), dc)),
                                   (float)__dadd_rn(cy, __ddiv_rn(__dmul_rn((double)sy, isx), dc)),
                                   (float)__dadd_rn(cz, __ddiv_rn(__dmul_rn((double)sz, isx), dc)),
                                   (float)__ddiv_rn(__dmul_rn((double)sw, isw), dc));
        counts[r] = (int)(j1 - j0);
    }
}

}  // namespace lidar

using namespace lidar;

extern "C" {

size_t lidar_voxel_sorted_workspace_bytes(int64_t n) {
    if (n < 0) return 0;
    return sort_layout(n).total;
}

int lidar_voxel_downsample_sorted(const void* d_points, int64_t n, double voxel_size, const double* h_origin3,
                                  const double* h_roi_lo3, const double* h_roi_hi3, int64_t* d_voxel_key, int32_t* d_inverse,
                                  float* d_centroids4, int32_t* d_counts, int64_t* d_unique_keys, lidar_sorted_desc* d_desc,
                                  void* d_ws, size_t ws_bytes, void* stream) {
    LIDAR_REQUIRE(n >= 0 && n < (1ll << 32) - 1, LIDAR_ERR_CAPACITY, "lidar_voxel_downsample_sorted: n must be < 2^32 - 1");
    LIDAR_REQUIRE(voxel_size > 0.0 && voxel_size == voxel_size, LIDAR_ERR_INVALID, "lidar_voxel_downsample_sorted: voxel_size must be > 0");
    LIDAR_REQUIRE(d_desc && (n == 0 || (d_points && d_voxel_key && d_inverse && d_centroids4 && d_counts && d_unique_keys)),
                  LIDAR_ERR_INVALID, "lidar_voxel_downsample_sorted: NULL argument");
    LIDAR_REQUIRE((h_roi_lo3 == nullptr) == (h_roi_hi3 == nullptr), LIDAR_ERR_INVALID,
                  "lidar_voxel_downsample_sorted: give both ROI corners or neither");
    const SortLayout L = sort_layout(n);
    LIDAR_REQUIRE(d_ws && ws_bytes >= L.total, LIDAR_ERR_WORKSPACE, "lidar_voxel_downsample_sorted: workspace too small (%zu < %zu)",
                  ws_bytes, L.total);
    cudaStream_t st = as_stream(stream);
    char* ws = static_cast<char*>(d_ws);
    SortWs* W = reinterpret_cast<SortWs*>(ws + L.off_ws);
    unsigned long long* tile_desc = reinterpret_cast<unsigned long long*>(ws + L.off_desc_tiles);
    unsigned long long* ka = reinterpret_cast<unsigned long long*>(ws + L.off_keys_a);
    unsigned long long* kb = reinterpret_cast<unsigned long long*>(ws + L.off_keys_b);
    unsigned* ia = reinterpret_cast<unsigned*>(ws + L.off_idx_a);
    unsigned* ib = reinterpret_cast<unsigned*>(ws + L.off_idx_b);
    unsigned* hist = reinterpret_cast<unsigned*>(ws + L.off_hist);
    unsigned* start = reinterpret_cast<unsigned*>(ws + L.off_start);
    LIDAR_CUDA_TRY(cudaMemsetAsync(W, 0, sizeof(SortWs), st));
    LIDAR_CUDA_TRY(cudaMemsetAsync(tile_desc, 0, sizeof(unsigned long long) * (size_t)L.tiles, st));
    LIDAR_CUDA_TRY(cudaMemsetAsync(d_desc, 0, sizeof(lidar_sorted_desc), st));
    SortParams P{};
    P.pts = static_cast<const float4*>(d_points);
    P.n = n;
    P.voxel = voxel_size;
    P.has_origin = h_origin3 != nullptr;
    for (int c = 0; c < 3; ++c) P.origin[c] = h_origin3 ? h_origin3[c] : 0.0;
    P.has_roi = h_roi_lo3 != nullptr;
    for (int c = 0; c < 3; ++c) {
        // B.2: fp32 compares against the bounds cast to fp32
        P.lo[c] = h_roi_lo3 ? (float)h_roi_lo3[c] : 0.f;
        P.hi[c] = h_roi_hi3 ? (float)h_roi_hi3[c] : 0.f;
    }
    int lg = 0;
    while ((1ll << lg) < (n > 1 ? n : 1)) ++lg;
    int grid = (int)((n + kSortThreads * 8 - 1) / (kSortThreads * 8));
    grid = grid < 1 ? 1 : (grid > 1024 ? 1024 : grid);
    k_sorted_bbox<<<grid, kSortThreads, 0, st>>>(P, W, d_desc, 62 - lg);
    LIDAR_CHECK_LAUNCH();
    if (n == 0) return LIDAR_OK;
    const int tiles = (int)L.tiles;
    k_sorted_keys<<<sm_count() * 4, kSortThreads, 0, st>>>(P, d_desc, ka, ia, reinterpret_cast<long long*>(d_voxel_key));
    LIDAR_CHECK_LAUNCH();
    for (int pass = 0; pass < 8; ++pass) {
        k_sort_hist<<<tiles, kSortThreads, 0, st>>>(pass, n, d_desc, ka, kb, ia, ib, hist, tiles);
        k_sort_scan<<<1, 1024, 0, st>>>(pass, d_desc, hist, (int64_t)256 * tiles);
        k_sort_scatter<<<tiles, kSortThreads, 0, st>>>(pass, n, d_desc, ka, kb, ia, ib, hist, tiles);
    }
    LIDAR_CHECK_LAUNCH();
    k_sorted_rank<<<tiles, kSortThreads, 0, st>>>(n, d_desc, ka, kb, ia, ib, tile_desc, &W->ticket, start,
                                                  reinterpret_cast<long long*>(d_unique_keys), d_inverse);
    LIDAR_CHECK_LAUNCH();
    k_sorted_reduce<<<sm_count() * 4, kSortThreads, 0, st>>>(P.pts, d_desc, ia, ib, start, reinterpret_cast<long long*>(d_unique_keys),
                                                             reinterpret_cast<float4*>(d_centroids4), d_counts);
    LIDAR_CHECK_LAUNCH();
    return LIDAR_OK;
}

}  // extern "C"
