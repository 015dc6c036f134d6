// K5 (+K6 fused) — one LiDAR frame: bbox -> voxel downsample (+ ground-plane density histogram).
//
// Voxel downsample is a NEW op (absent from the reference; contract = SURVEY.md Appendix B.1); the
// fused density grid has calculate_grid_density semantics (utils/data_processing.py:282-328):
// margin 2g, np.arange edges (fill rule e(i) = a + i*fl(fl(a+g)-a), Appendix A.2), histogramdd
// binning (Appendix A.1), counts indexed [x][y].
//
// Design: "occupancy-bitmap ranking" instead of a key sort.
//   The voxel key space (Dx*Dy*Dz cells) is held as a bitmap (1 bit per cell, 20 MB for a
//   100 m x 100 m x 2 m frame at 0.05 m — L2 resident on B200).  The rank of a voxel in ascending
//   key order is the number of occupied cells before it, i.e. a popcount prefix over the bitmap.
//   That yields exactly the output order of a stable sort-by-key + segmented reduce, without
//   moving a single point, and every step is order independent:
//     k_frame_prep      min/max of x,y,z,intensity; zeroes what the previous frame dirtied (bitmap,
//                       duplicate filter, grid); the last CTA derives the frame descriptor (origin,
//                       dims, key space, fixed-point scales, histogram edges) ON DEVICE
//     k_frame_mark      per point: voxel key -> d_voxel_key, atomicOr into the bitmap (a second hit
//                       on the same bit flags the voxel in a 512 KB hashed duplicate filter),
//                       density bin -> RED.ADD into the grid              (reads 16 B, writes 4 B)
//     k_frame_scan      one-wave scan of bitmap popcounts -> prefix per 256-bit group
//     k_frame_rank      per point: rank = prefix + popc(bits below) -> d_inverse.  Voxels not in
//                       the duplicate filter hold exactly one point (94 % of a crowd frame): the
//                       point IS the centroid and is stored directly.  The rest accumulate
//                       (p - ref) in 2^-k fixed point with integer atomics and enlist the voxel.
//     k_frame_finalize  enlisted voxels only: centroid = ref + sum/count; re-zeroes accumulators
//   Integer accumulation makes centroids independent of the order in which atomics land:
//   bit-identical run to run, and (p - ref)*2^k is an exact integer, so the sums are exact.
//
// Nothing in a frame needs the host: capacities are fixed per stream of frames, the descriptor is
// read back together with the results.
#include "common.cuh"

namespace lidar {

constexpr int kFrameThreads = 256;
// Occupancy groups: one 32-byte sector = {prefix, 7 x 32 occupancy bits} = 224 voxels.  Packing the
// popcount prefix next to the bits means the rank of a point costs ONE random sector read.
constexpr int kGroupVoxels = 224;
constexpr int kScanGroupsPerThread = 4;                               // each thread scans 4 adjacent groups
constexpr int kScanTileGroups = kFrameThreads * kScanGroupsPerThread; // 1024 groups = 229 376 voxels per tile

constexpr int kFilterBits = 22;                                       // duplicate filter: 2^22 bits = 512 KB
constexpr int kFilterWords = 1 << (kFilterBits - 5);

struct FrameWsLayout {
    size_t off_partial, off_ctrl, off_groups, off_tile_desc, off_acc, off_cnt, off_filter, off_cta_desc, total;
    int64_t groups, tiles;
};

struct FrameCtrl {
    unsigned int bbox_ticket;
    unsigned int scan_ticket;
    unsigned int grid_bar;     // fused kernel: grid-barrier arrival counter (reset by the last CTA to exit)
    unsigned int exit_ticket;  // fused kernel: CTAs that have passed the last barrier
    unsigned long long dirty_groups;  // occupancy groups the previous frame may have set
};

constexpr int kFusedMaxCtas = 1024;   // cap of the fused kernel's grid (one slot of cta_desc each)

__device__ __forceinline__ unsigned filter_slot(int key) { return ((unsigned)key * 2654435761u) >> (32 - kFilterBits); }

constexpr int kBboxMaxBlocks = 1024;

static FrameWsLayout frame_layout(const lidar_frame_caps& c) {
    FrameWsLayout L;
    L.groups = (c.max_key_space + kGroupVoxels - 1) / kGroupVoxels;
    // round up to whole tiles so the scan never needs a ragged tail
    L.tiles = (L.groups + kScanTileGroups - 1) / kScanTileGroups;
    L.groups = L.tiles * kScanTileGroups;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t a = ws_align(o); o = a + bytes; return a; };
    L.off_partial = take(sizeof(double) * 8 * kBboxMaxBlocks);
    L.off_ctrl = take(sizeof(FrameCtrl));
    L.off_groups = take(sizeof(uint32_t) * 8 * L.groups);
    L.off_tile_desc = take(sizeof(unsigned long long) * L.tiles);
    L.off_acc = take(sizeof(long long) * 4 * c.max_points);
    L.off_cnt = take(sizeof(int32_t) * c.max_points);
    L.off_filter = take(sizeof(uint32_t) * kFilterWords);
    L.off_cta_desc = take(sizeof(unsigned long long) * kFusedMaxCtas);
    L.total = ws_align(o);
    return L;
}

struct FrameParams {
    const float4* pts;
    int64_t n;
    double voxel, grid;
    double origin[3];
    double xyr[4];
    int has_origin, has_range;
    int64_t max_key_space;
    int max_nx, max_ny;
    int fix_bits_budget;  // 62 - ceil(log2(n))
};

// ---- descriptor derivation (one thread) -------------------------------------------------------
__device__ void derive_desc(const FrameParams& P, const double* bb, lidar_frame_desc* D) {
    int status = 0;
    for (int c = 0; c < 4; ++c) {
        D->bbox_min[c] = bb[c];
        D->bbox_max[c] = bb[4 + c];
    }
    D->voxel = P.voxel;
    D->n_points = P.n;
    D->n_voxels = 0;
    long long ks = 1;
    for (int c = 0; c < 3; ++c) {
        const double o = P.has_origin ? P.origin[c] : bb[c];
        D->origin[c] = o;
        if (bb[c] < o) status = LIDAR_ERR_INVALID;  // a point below the origin would index < 0
        const double span = floor(__ddiv_rn(__dsub_rn(bb[4 + c], o), P.voxel));
        long long d = (span >= 0.0 && span < 2147483000.0) ? (long long)span + 1 : 0;
        if (d <= 0) { d = 1; if (P.n > 0) status = status ? status : LIDAR_ERR_CAPACITY; }
        D->dims[c] = (int)d;
        // overflow-safe product
        if (ks > (1ll << 40)) status = status ? status : LIDAR_ERR_CAPACITY;
        ks *= d;
    }
    D->origin[3] = 0.0;
    D->dims[3] = 0;
    D->key_space = ks;
    if (ks > P.max_key_space || ks >= (1ll << 31)) status = status ? status : LIDAR_ERR_CAPACITY;
    // fixed point: |p - ref| < 2*voxel + 2^-24, n members at most  ->  that * 2^k * n < 2^62
    {
        int e;
        frexp(P.voxel * 2.0 + 0x1p-24, &e);  // bound < 2^e
        D->fix_scale_xyz = ldexp(1.0, P.fix_bits_budget - e);
        double wmax = fmax(fabs(bb[3]), fabs(bb[7]));
        if (!(wmax > 0.0) || !isfinite(wmax)) wmax = 1.0;
        frexp(wmax, &e);
        D->fix_scale_w = ldexp(1.0, P.fix_bits_budget - e - 1);
    }
    // density grid edges, numpy arange rule (data_processing.py:305-313)
    D->grid = P.grid;
    D->nx = D->ny = 0;
    D->ex0 = D->ex1 = D->exd = D->ey0 = D->ey1 = D->eyd = 0.0;
    if (P.grid > 0.0) {
        const double g = P.grid;
        const double margin = __dmul_rn(g, 2.0);
        double lo[2], hi[2];
        lo[0] = P.has_range ? P.xyr[0] : bb[0];
        hi[0] = P.has_range ? P.xyr[1] : bb[4];
        lo[1] = P.has_range ? P.xyr[2] : bb[1];
        hi[1] = P.has_range ? P.xyr[3] : bb[5];
        for (int c = 0; c < 2; ++c) {
            const double a = __dsub_rn(lo[c], margin);
            const double stop = __dadd_rn(__dadd_rn(hi[c], margin), g);
            const double len = ceil(__ddiv_rn(__dsub_rn(stop, a), g));
            int nedges = (len > 0.0 && len < 1.0e9) ? (int)len : 0;
            const double e1 = __dadd_rn(a, g);
            const double delta = __dsub_rn(e1, a);
            int nb = nedges - 1;
            if (nb < 1) { nb = 0; if (P.n > 0) status = status ? status : LIDAR_ERR_CAPACITY; }
            if (c == 0) { D->ex0 = a; D->ex1 = e1; D->exd = delta; D->nx = nb; }
            else        { D->ey0 = a; D->ey1 = e1; D->eyd = delta; D->ny = nb; }
        }
        if (D->nx > P.max_nx || D->ny > P.max_ny) status = status ? status : LIDAR_ERR_CAPACITY;
    }
    if (P.n == 0) status = 0;
    D->status = status;
    // the fp32 index guess needs an origin that fp32 represents exactly (true for a bbox-derived origin)
    int fast = 1;
    for (int c = 0; c < 3; ++c)
        if ((double)(float)D->origin[c] != D->origin[c]) fast = 0;
    if ((double)(float)P.voxel == 0.0) fast = 0;
    D->fast_f32 = fast;
    // Granlund-Montgomery: for 0 <= n < 2^31 and 2^(l-1) < d <= 2^l,  n / d == (n * ceil(2^(31+l)/d)) >> (31+l)
    for (int c = 0; c < 2; ++c) {
        const unsigned long long d = (unsigned long long)(c == 0 ? D->dims[2] : D->dims[1]);
        int l = 0;
        while ((1ull << l) < d) ++l;
        const unsigned long long num = 1ull << (31 + l);
        const unsigned m = (unsigned)((num + d - 1) / d);
        if (c == 0) { D->magic_dz = m; D->shift_dz = l; } else { D->magic_dy = m; D->shift_dy = l; }
    }
}

__device__ __forceinline__ int magic_div(int n, unsigned magic, int shift) {
    return (int)(((unsigned long long)(unsigned)n * magic) >> (31 + shift));
}
// key -> (ix, iy, iz) without integer division instructions
__device__ __forceinline__ void decode_key(int key, const lidar_frame_desc& D, int& ix, int& iy, int& iz) {
    const int t = magic_div(key, D.magic_dz, D.shift_dz);
    iz = key - t * D.dims[2];
    ix = magic_div(t, D.magic_dy, D.shift_dy);
    iy = t - ix * D.dims[1];
}

// ---- k_frame_prep -----------------------------------------------------------------------------
__global__ void __launch_bounds__(kFrameThreads)
k_frame_prep(FrameParams P, double* __restrict__ partial, FrameCtrl* __restrict__ ctrl,
             lidar_frame_desc* __restrict__ D, int32_t* __restrict__ grid_out, int grid_cap,
             unsigned long long* __restrict__ tile_desc, int64_t tiles, uint32_t* __restrict__ groups,
             int64_t groups_cap, uint32_t* __restrict__ filter) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    // issue this thread's point loads first so they are in flight while the zeroing stores drain
    LoadF32x4 L{P.pts};
    float mn[4] = {INFINITY, INFINITY, INFINITY, INFINITY};
    float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    auto fold = [&](const float4& v) {
        mn[0] = fminf(mn[0], v.x); mx[0] = fmaxf(mx[0], v.x);
        mn[1] = fminf(mn[1], v.y); mx[1] = fmaxf(mx[1], v.y);
        mn[2] = fminf(mn[2], v.z); mx[2] = fmaxf(mx[2], v.z);
        mn[3] = fminf(mn[3], v.w); mx[3] = fmaxf(mx[3], v.w);
    };
    int64_t i = t0;
    for (; i + 3 * stride < P.n; i += 4 * stride) {
        const float4 a = L.raw(i), b = L.raw(i + stride), c = L.raw(i + 2 * stride), d = L.raw(i + 3 * stride);
        fold(a); fold(b); fold(c); fold(d);
    }
    for (; i < P.n; i += stride) fold(L.raw(i));
    // zero what the previous frame dirtied and what later kernels accumulate into
    {
        int64_t dirty = (int64_t)ctrl->dirty_groups;
        if (dirty > groups_cap) dirty = groups_cap;
        uint4* b4 = reinterpret_cast<uint4*>(groups);
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        for (int64_t k = t0; k < dirty * 2; k += stride) b4[k] = z;
        uint4* f4 = reinterpret_cast<uint4*>(filter);
        for (int64_t k = t0; k < kFilterWords / 4; k += stride) f4[k] = z;
        for (int64_t k = t0; k < grid_cap; k += stride) grid_out[k] = 0;
        for (int64_t k = t0; k < tiles; k += stride) tile_desc[k] = 0ull;
    }
    __shared__ float s_v[kFrameThreads / 32][8];
    __shared__ bool s_last;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o));
            mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
        }
    }
    const int warp = threadIdx.x >> 5;
    if (lane_id() == 0) {
#pragma unroll
        for (int c = 0; c < 4; ++c) { s_v[warp][c] = mn[c]; s_v[warp][4 + c] = mx[c]; }
    }
    __syncthreads();
    if (threadIdx.x < 8) {
        const bool is_max = threadIdx.x >= 4;
        float v = is_max ? -INFINITY : INFINITY;
        for (int w = 0; w < kFrameThreads / 32; ++w) v = is_max ? fmaxf(v, s_v[w][threadIdx.x]) : fminf(v, s_v[w][threadIdx.x]);
        partial[(size_t)blockIdx.x * 8 + threadIdx.x] = (double)v;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&ctrl->bbox_ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    __shared__ double s_bb[8];
    {
        // all 256 threads fold the per-CTA partials: thread t owns channel t&7 of CTAs t>>3, +32, ...
        const int ch = threadIdx.x & 7;
        const bool is_max = ch >= 4;
        double v = is_max ? -INFINITY : INFINITY;
        for (unsigned b = threadIdx.x >> 3; b < gridDim.x; b += kFrameThreads / 8) {
            const double q = __ldcg(partial + (size_t)b * 8 + ch);
            v = is_max ? fmax(v, q) : fmin(v, q);
        }
        // lanes with equal (lane & 7) hold the same channel: xor-shuffle over 8 and 16
        for (int o = 8; o < 32; o <<= 1) {
            const double q = __shfl_xor_sync(0xffffffffu, v, o);
            v = is_max ? fmax(v, q) : fmin(v, q);
        }
        __shared__ double s_part[kFrameThreads / 32][8];
        if (lane_id() < 8) s_part[warp][lane_id()] = v;
        __syncthreads();
        if (threadIdx.x < 8) {
            double r = s_part[0][threadIdx.x];
            for (int w = 1; w < kFrameThreads / 32; ++w)
                r = (threadIdx.x >= 4) ? fmax(r, s_part[w][threadIdx.x]) : fmin(r, s_part[w][threadIdx.x]);
            s_bb[threadIdx.x] = r;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        derive_desc(P, s_bb, D);
        ctrl->bbox_ticket = 0u;
        ctrl->scan_ticket = 0u;
        // groups this frame may set (whole scan tiles), remembered for the next frame's zeroing
        const long long ng = (D->key_space + kGroupVoxels - 1) / kGroupVoxels;
        const long long tiles_used = (ng + kScanTileGroups - 1) / kScanTileGroups;
        ctrl->dirty_groups = D->status == 0 && P.n > 0 ? (unsigned long long)(tiles_used * kScanTileGroups) : 0ull;
    }
}

// floor(fl(d / v)) without the DDIV sequence on the fast path: d * fl(1/v) is within 1.5 ulp of d/v,
// so unless it lands within 4 ulp of an integer its floor equals the floor of the correctly rounded
// quotient; the (astronomically rare) near-integer case takes the real division.
__device__ __forceinline__ int floor_div_exact(double d, double v, double rinv) {
    const double qh = __dmul_rn(d, rinv);
    const double fl = floor(qh);
    const double fr = __dsub_rn(qh, fl);
    const double tol = __dadd_rn(__dmul_rn(fabs(qh), 0x1p-50), 1e-300);
    if (fr > tol && __dsub_rn(1.0, fr) > tol) return (int)fl;
    return (int)floor(__ddiv_rn(d, v));
}

// Accumulation reference of a voxel: its corner snapped to a multiple of 2^-24 m.  p is an fp32 value
// and ref a multiple of 2^-24, so (p - ref) is exact in fp64 and (p - ref) * 2^k is an exact integer
// for every |p| >= 2^-(k-23): the integer sums below are then the EXACT sums of the members.
__device__ __forceinline__ double voxel_ref(double o, int i, double v) {
    const double c = __dadd_rn(o, __dmul_rn((double)i, v));
    return __dmul_rn(nearbyint(__dmul_rn(c, 16777216.0)), 1.0 / 16777216.0);
}

// analytic arange edge (DOUBLE_fill): e(0)=a, e(1)=fl(a+g), e(i)=fl(a + fl(i*delta))
__device__ __forceinline__ double arange_edge(double a, double e1, double d, int i) {
    return i == 0 ? a : (i == 1 ? e1 : __dadd_rn(a, __dmul_rn((double)i, d)));
}
__device__ __forceinline__ int arange_bin(double x, double a, double e1, double d, double rd, int nb) {
    const double hi = arange_edge(a, e1, d, nb);
    if (!(x >= a) || !(x <= hi)) return -1;
    if (x == hi) return nb - 1;
    int k = (int)floor(__dmul_rn(__dsub_rn(x, a), rd));   // guess; corrected against the exact edges
    k = k < 0 ? 0 : (k > nb - 1 ? nb - 1 : k);
    while (x < arange_edge(a, e1, d, k)) --k;
    while (x >= arange_edge(a, e1, d, k + 1)) ++k;
    return k;
}

// fp32 guess of floor((p - o) / v), accepted only when the fractional part is farther from an integer
// than the worst-case fp32 error; returns false when the exact fp64 path must decide
__device__ __forceinline__ bool fast_voxel_index(float p, float of, float rvf, int& k) {
    const float q = __fmul_rn(__fsub_rn(p, of), rvf);
    const float kf = floorf(q);
    const float fr = __fsub_rn(q, kf);
    const float eps = fmaf(q, 4.0e-7f, 2.0e-6f);     // 3 roundings of 2^-24 each, with margin
    k = (int)kf;
    return fr > eps && fr < 1.0f - eps && q < 1.0e6f;
}
// fp32 guess of the histogram bin, VERIFIED against the exact fp64 edges; falls back to arange_bin
__device__ __forceinline__ int fast_arange_bin(float xf, double x, float af, float rdf, double a, double e1, double d,
                                               double rd, int nb) {
    int k = (int)floorf(__fmul_rn(__fsub_rn(xf, af), rdf));
    k = k < 0 ? 0 : (k > nb - 1 ? nb - 1 : k);
    double lo = __dadd_rn(a, __dmul_rn((double)k, d));
    double hi = __dadd_rn(a, __dmul_rn((double)(k + 1), d));
    lo = k == 1 ? e1 : lo;
    hi = k == 0 ? e1 : hi;
    if (x >= lo && x < hi) return k;
    return arange_bin(x, a, e1, d, rd, nb);
}

// streaming accesses: every point / key / inverse entry is touched once per pass, keep them from
// evicting the L2-resident working set (occupancy groups, voxel records)
__device__ __forceinline__ unsigned long long evict_first_policy() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ int ld_stream_s32(const int* p) {
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;"
                 : "=r"(v) : "l"(p), "l"(evict_first_policy()));
    return v;
}
__device__ __forceinline__ void st_stream_s32(int* p, int v) {
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.s32 [%0], %1, %2;"
                 ::"l"(p), "r"(v), "l"(evict_first_policy()) : "memory");
}

// ---- k_frame_mark -----------------------------------------------------------------------------
struct MarkConst {
    float of[3], rvf, axf, ayf, rdxf, rdyf;
    double rv, rdx, rdy;
    int fast;
};

__device__ __forceinline__ void mark_point(const float4& q, const lidar_frame_desc& D, const MarkConst& K,
                                           bool do_grid, int64_t i, int32_t* __restrict__ voxel_key,
                                           uint32_t* __restrict__ groups, uint32_t* __restrict__ filter,
                                           int32_t* __restrict__ grid_out) {
    int ix, iy, iz;
    const bool fx = K.fast && fast_voxel_index(q.x, K.of[0], K.rvf, ix);
    const bool fy = K.fast && fast_voxel_index(q.y, K.of[1], K.rvf, iy);
    const bool fz = K.fast && fast_voxel_index(q.z, K.of[2], K.rvf, iz);
    if (!fx) ix = floor_div_exact(__dsub_rn((double)q.x, D.origin[0]), D.voxel, K.rv);
    if (!fy) iy = floor_div_exact(__dsub_rn((double)q.y, D.origin[1]), D.voxel, K.rv);
    if (!fz) iz = floor_div_exact(__dsub_rn((double)q.z, D.origin[2]), D.voxel, K.rv);
    const int key = (ix * D.dims[1] + iy) * D.dims[2] + iz;
    st_stream_s32(voxel_key + i, key);
    const unsigned g = (unsigned)key / kGroupVoxels, b = (unsigned)key - g * kGroupVoxels;
    const unsigned bit = 1u << (b & 31);
    const unsigned old = atomicOr(&groups[(size_t)g * 8 + 1 + (b >> 5)], bit);
    if (do_grid) {
        const int bx = fast_arange_bin(q.x, (double)q.x, K.axf, K.rdxf, D.ex0, D.ex1, D.exd, K.rdx, D.nx);
        const int by = fast_arange_bin(q.y, (double)q.y, K.ayf, K.rdyf, D.ey0, D.ey1, D.eyd, K.rdy, D.ny);
        if (bx >= 0 && by >= 0) atomicAdd(&grid_out[bx * D.ny + by], 1);
    }
    if (old & bit) {   // the voxel already had a point: flag it as multi-member
        const unsigned h = filter_slot(key);
        atomicOr(&filter[h >> 5], 1u << (h & 31));
    }
}

__global__ void __launch_bounds__(kFrameThreads)
k_frame_mark(const float4* __restrict__ pts, const lidar_frame_desc* __restrict__ Dg,
             int32_t* __restrict__ voxel_key, uint32_t* __restrict__ groups, uint32_t* __restrict__ filter,
             int32_t* __restrict__ grid_out) {
    __shared__ lidar_frame_desc D;
    LoadF32x4 L{pts};
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (threadIdx.x == 0) D = *Dg;
    __syncthreads();
    if (D.status != 0) return;
    const int64_t n = D.n_points;
    const bool do_grid = D.grid > 0.0;
    MarkConst K;
    K.rv = __ddiv_rn(1.0, D.voxel);
    K.rdx = do_grid ? __ddiv_rn(1.0, D.exd) : 0.0;
    K.rdy = do_grid ? __ddiv_rn(1.0, D.eyd) : 0.0;
    K.of[0] = (float)D.origin[0]; K.of[1] = (float)D.origin[1]; K.of[2] = (float)D.origin[2];
    K.rvf = (float)K.rv;
    K.axf = (float)D.ex0; K.ayf = (float)D.ey0; K.rdxf = (float)K.rdx; K.rdyf = (float)K.rdy;
    K.fast = D.fast_f32;
    int64_t i = t0;
    for (; i + stride < n; i += 2 * stride) {
        const float4 a = L.raw(i), b = L.raw(i + stride);
        mark_point(a, D, K, do_grid, i, voxel_key, groups, filter, grid_out);
        mark_point(b, D, K, do_grid, i + stride, voxel_key, groups, filter, grid_out);
    }
    if (i < n) mark_point(L.raw(i), D, K, do_grid, i, voxel_key, groups, filter, grid_out);
}

// ---- k_frame_scan -----------------------------------------------------------------------------
// All tiles of a frame run in ONE wave, so the classic chained look-back would degenerate into a
// serial chain.  Instead every CTA publishes its tile total and then sums the totals of ALL earlier
// tiles with its 256 threads (spinning only on totals not yet published).  The exclusive prefix of
// every group is stored in word 0 of the group itself.
__global__ void __launch_bounds__(kFrameThreads)
k_frame_scan(uint32_t* __restrict__ groups, unsigned long long* __restrict__ tile_desc,
             FrameCtrl* __restrict__ ctrl, lidar_frame_desc* __restrict__ Dg) {
    __shared__ int s_tile;
    __shared__ unsigned s_warp_sum[kFrameThreads / 32];
    __shared__ unsigned long long s_look[kFrameThreads / 32];
    if (Dg->status != 0) return;
    const int64_t ng = (Dg->key_space + kGroupVoxels - 1) / kGroupVoxels;
    const int n_tiles = (int)((ng + kScanTileGroups - 1) / kScanTileGroups);
    const unsigned lane = lane_id();
    const int warp = threadIdx.x >> 5;
    while (true) {
        if (threadIdx.x == 0) s_tile = (int)atomicAdd(&ctrl->scan_ticket, 1u);
        __syncthreads();
        const int tile = s_tile;
        if (tile >= n_tiles) break;
        // the group array is padded to whole tiles, so the loads never run off the end
        uint32_t* base = groups + ((size_t)tile * kScanTileGroups + (size_t)threadIdx.x * kScanGroupsPerThread) * 8;
        const uint4* src = reinterpret_cast<const uint4*>(base);
        unsigned gcnt[kScanGroupsPerThread];
        unsigned cnt = 0;
        uint4 w[2 * kScanGroupsPerThread];
#pragma unroll
        for (int g = 0; g < 2 * kScanGroupsPerThread; ++g) w[g] = src[g];
#pragma unroll
        for (int g = 0; g < kScanGroupsPerThread; ++g) {
            const uint4 a = w[2 * g], b = w[2 * g + 1];   // a.x is the (still zero) prefix slot
            gcnt[g] = __popc(a.y) + __popc(a.z) + __popc(a.w) + __popc(b.x) + __popc(b.y) + __popc(b.z) + __popc(b.w);
            cnt += gcnt[g];
        }
        unsigned inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= (unsigned)o) inc += t;
        }
        if (lane == 31) s_warp_sum[warp] = inc;
        __syncthreads();
        unsigned warp_off = 0, total = 0;
#pragma unroll
        for (int w2 = 0; w2 < kFrameThreads / 32; ++w2) {
            const unsigned sv = s_warp_sum[w2];
            if (w2 < warp) warp_off += sv;
            total += sv;
        }
        if (threadIdx.x == 0) st_relaxed_u64(tile_desc + tile, kScanAgg | (unsigned long long)total);
        unsigned long long look = 0ull;
        for (int t = threadIdx.x; t < tile; t += kFrameThreads) {
            unsigned long long d;
            do { d = ld_relaxed_u64(tile_desc + t); } while ((d >> 62) == 0ull);
            look += d & kScanValMask;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) look += __shfl_xor_sync(0xffffffffu, look, o);
        if (lane == 0) s_look[warp] = look;
        __syncthreads();
        unsigned long long excl = 0ull;
#pragma unroll
        for (int w2 = 0; w2 < kFrameThreads / 32; ++w2) excl += s_look[w2];
        if (threadIdx.x == 0 && tile == n_tiles - 1) Dg->n_voxels = (int64_t)(excl + total);
        unsigned run = (unsigned)excl + warp_off + (inc - cnt);
#pragma unroll
        for (int g = 0; g < kScanGroupsPerThread; ++g) {
            base[g * 8] = run;
            run += gcnt[g];
        }
        __syncthreads();
    }
}

// ---- k_frame_rank -----------------------------------------------------------------------------
// The accumulate path is needed by ~12 % of the points but, taken in place, would be executed by almost
// every warp (divergence).  Each warp therefore parks those points in a private shared-memory ring and
// drains it 32 at a time with all lanes active.
struct MultiItem {
    float4 q;
    int key;
    unsigned r;
    int pad[2];
};
constexpr int kRingSize = 64;

__device__ __forceinline__ void accumulate_multi(const MultiItem& it, const lidar_frame_desc& D,
                                                 long long* __restrict__ acc, int32_t* __restrict__ cnt,
                                                 lidar_voxel* __restrict__ voxels) {
    int ix, iy, iz;
    decode_key(it.key, D, ix, iy, iz);
    const double cx = voxel_ref(D.origin[0], ix, D.voxel);
    const double cy = voxel_ref(D.origin[1], iy, D.voxel);
    const double cz = voxel_ref(D.origin[2], iz, D.voxel);
    const long long fx = __double2ll_rn(__dmul_rn(__dsub_rn((double)it.q.x, cx), D.fix_scale_xyz));
    const long long fy = __double2ll_rn(__dmul_rn(__dsub_rn((double)it.q.y, cy), D.fix_scale_xyz));
    const long long fz = __double2ll_rn(__dmul_rn(__dsub_rn((double)it.q.z, cz), D.fix_scale_xyz));
    const long long fw = __double2ll_rn(__dmul_rn((double)it.q.w, D.fix_scale_w));
    unsigned long long* A = reinterpret_cast<unsigned long long*>(acc + (size_t)it.r * 4);
    // results unused: these compile to RED (fire and forget), nothing in the warp waits on them
    atomicAdd(A + 0, (unsigned long long)fx);
    atomicAdd(A + 1, (unsigned long long)fy);
    atomicAdd(A + 2, (unsigned long long)fz);
    atomicAdd(A + 3, (unsigned long long)fw);
    atomicAdd(cnt + it.r, 1);
    voxels[it.r].key = it.key;   // every member stores the same value
}

// returns the rank and whether the voxel is (possibly) multi-member
__device__ __forceinline__ unsigned rank_of(int key, const uint32_t* __restrict__ groups) {
    const unsigned g = (unsigned)key / kGroupVoxels, b = (unsigned)key - g * kGroupVoxels;
    const uint4* gw = reinterpret_cast<const uint4*>(groups + (size_t)g * 8);
    const uint4 a4 = gw[0], b4 = gw[1];
    const unsigned w[8] = {a4.x, a4.y, a4.z, a4.w, b4.x, b4.y, b4.z, b4.w};
    const int wi = 1 + (int)(b >> 5);
    unsigned r = w[0];
#pragma unroll
    for (int k = 1; k < 8; ++k) {
        if (k < wi) r += __popc(w[k]);
        else if (k == wi) r += __popc(w[k] & ((1u << (b & 31)) - 1u));
    }
    return r;
}

__global__ void __launch_bounds__(kFrameThreads)
k_frame_rank(const float4* __restrict__ pts, const lidar_frame_desc* __restrict__ Dg,
             const int32_t* __restrict__ voxel_key, const uint32_t* __restrict__ groups,
             const uint32_t* __restrict__ filter, int32_t* __restrict__ inverse, long long* __restrict__ acc,
             int32_t* __restrict__ cnt, lidar_voxel* __restrict__ voxels) {
    __shared__ lidar_frame_desc D;
    __shared__ MultiItem s_ring[kFrameThreads / 32][kRingSize];
    if (threadIdx.x == 0) D = *Dg;
    __syncthreads();
    if (D.status != 0) return;
    const int64_t n = D.n_points;
    LoadF32x4 L{pts};
    const unsigned lane = lane_id();
    MultiItem* ring = s_ring[threadIdx.x >> 5];
    unsigned head = 0, count = 0;     // warp-uniform
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t n_round = ((n + 31) / 32) * 32;   // keep whole warps in the loop for the ballots
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
        const bool live = i < n;
        bool multi = false;
        MultiItem it;
        if (live) {
            it.key = ld_stream_s32(voxel_key + i);
            it.q = L.raw(i);
            it.r = rank_of(it.key, groups);
            st_stream_s32(inverse + i, (int)it.r);
            const unsigned h = filter_slot(it.key);
            multi = (__ldg(filter + (h >> 5)) >> (h & 31)) & 1u;
            if (!multi) {
                // exactly one point in this voxel: it is the centroid.  One full 32-byte sector store.
                float4* rec = reinterpret_cast<float4*>(voxels + it.r);
                rec[0] = it.q;
                rec[1] = make_float4(__int_as_float(1), __int_as_float(it.key), 0.f, 0.f);
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, multi);
        if (m) {
            if (multi) ring[(head + count + __popc(m & lanemask_lt())) % kRingSize] = it;
            count += __popc(m);
            __syncwarp();
            if (count >= 32) {
                accumulate_multi(ring[(head + lane) % kRingSize], D, acc, cnt, voxels);
                head = (head + 32) % kRingSize;
                count -= 32;
                __syncwarp();
            }
        }
    }
    if (lane < count) accumulate_multi(ring[(head + lane) % kRingSize], D, acc, cnt, voxels);
}

// ---- k_frame_finalize -------------------------------------------------------------------------
// cnt[r] != 0 exactly for the voxels that went through the accumulators (multi-member voxels and the
// few singletons that collided in the duplicate filter): a coalesced sweep over cnt finds them; the
// ~6 % of hits are parked in a per-warp ring and finished 32 at a time, so the fp64 divisions run with
// full warps instead of being executed, mostly masked, by every warp.
__device__ __forceinline__ void finalize_voxel(unsigned r, const lidar_frame_desc& D, double isx, double isw,
                                               long long* __restrict__ acc, int32_t* __restrict__ cnt,
                                               lidar_voxel* __restrict__ voxels) {
    const int c = cnt[r];
    const int key = voxels[r].key;
    longlong2* A = reinterpret_cast<longlong2*>(acc + (size_t)r * 4);
    const longlong2 s01 = A[0], s23 = A[1];
    int ix, iy, iz;
    decode_key(key, D, ix, iy, iz);
    const double dc = (double)c;
    const double cx = voxel_ref(D.origin[0], ix, D.voxel);
    const double cy = voxel_ref(D.origin[1], iy, D.voxel);
    const double cz = voxel_ref(D.origin[2], iz, D.voxel);
    float4 o;
    o.x = (float)__dadd_rn(cx, __ddiv_rn(__dmul_rn((double)s01.x, isx), dc));
    o.y = (float)__dadd_rn(cy, __ddiv_rn(__dmul_rn((double)s01.y, isx), dc));
    o.z = (float)__dadd_rn(cz, __ddiv_rn(__dmul_rn((double)s23.x, isx), dc));
    o.w = (float)__ddiv_rn(__dmul_rn((double)s23.y, isw), dc);
    float4* rec = reinterpret_cast<float4*>(voxels + r);
    rec[0] = o;
    rec[1] = make_float4(__int_as_float(c), __int_as_float(key), 0.f, 0.f);
    // restore the all-zero invariant of the accumulators for the next frame
    A[0] = make_longlong2(0, 0);
    A[1] = make_longlong2(0, 0);
    cnt[r] = 0;
}

__global__ void __launch_bounds__(kFrameThreads)
k_frame_finalize(const lidar_frame_desc* __restrict__ Dg, long long* __restrict__ acc, int32_t* __restrict__ cnt,
                 lidar_voxel* __restrict__ voxels) {
    __shared__ lidar_frame_desc D;
    __shared__ unsigned s_ring[kFrameThreads / 32][kRingSize];
    if (threadIdx.x == 0) D = *Dg;
    __syncthreads();
    if (D.status != 0) return;
    const int64_t V = D.n_voxels;
    const double isx = 1.0 / D.fix_scale_xyz, isw = 1.0 / D.fix_scale_w;  // powers of two: exact
    const unsigned lane = lane_id();
    unsigned* ring = s_ring[threadIdx.x >> 5];
    unsigned head = 0, count = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t v_round = ((V + 31) / 32) * 32;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < v_round; r += stride) {
        const bool hit = r < V && __ldg(cnt + r) != 0;
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (m) {
            if (hit) ring[(head + count + __popc(m & lanemask_lt())) % kRingSize] = (unsigned)r;
            count += __popc(m);
            __syncwarp();
            if (count >= 32) {
                finalize_voxel(ring[(head + lane) % kRingSize], D, isx, isw, acc, cnt, voxels);
                head = (head + 32) % kRingSize;
                count -= 32;
                __syncwarp();
            }
        }
    }
    if (lane < count) finalize_voxel(ring[(head + lane) % kRingSize], D, isx, isw, acc, cnt, voxels);
}

static int g_ctas_per_sm = 8;   // grid cap of the per-point frame kernels, in CTAs per SM (tuning knob)

static int frame_grid(int64_t n, int per_thread) {
    int64_t want = (n + (int64_t)kFrameThreads * per_thread - 1) / ((int64_t)kFrameThreads * per_thread);
    if (want < 1) want = 1;
    const int64_t cap = (int64_t)sm_count() * g_ctas_per_sm;
    return (int)(want < cap ? want : cap);
}

}  // namespace lidar

using namespace lidar;

extern "C" {

int lidar_frame_set_ctas_per_sm(int ctas_per_sm) {
    LIDAR_REQUIRE(ctas_per_sm >= 1 && ctas_per_sm <= 8, LIDAR_ERR_INVALID, "lidar_frame_set_ctas_per_sm: 1..8");
    g_ctas_per_sm = ctas_per_sm;
    return LIDAR_OK;
}

size_t lidar_frame_workspace_bytes(const lidar_frame_caps* caps) {
    if (!caps || caps->max_points < 0 || caps->max_key_space <= 0) return 0;
    return frame_layout(*caps).total;
}

int lidar_frame_workspace_init(void* d_ws, size_t ws_bytes, const lidar_frame_caps* caps, void* stream) {
    LIDAR_REQUIRE(caps != nullptr, LIDAR_ERR_INVALID, "lidar_frame_workspace_init: caps is NULL");
    const FrameWsLayout L = frame_layout(*caps);
    LIDAR_REQUIRE(d_ws && ws_bytes >= L.total, LIDAR_ERR_WORKSPACE,
                  "lidar_frame_workspace_init: workspace too small (%zu < %zu)", ws_bytes, L.total);
    LIDAR_CUDA_TRY(cudaMemsetAsync(d_ws, 0, L.total, as_stream(stream)));
    return LIDAR_OK;
}

static int frame_voxel_density_impl(const void* d_points, int64_t n, double voxel_size, double grid_size,
                              const double* h_origin3, const double* h_xy_range4,
                              int32_t* d_voxel_key, int32_t* d_inverse, lidar_voxel* d_voxels, int32_t* d_grid,
                              lidar_frame_desc* d_desc, const lidar_frame_caps* caps, void* d_ws,
                              size_t ws_bytes, void* stream, void** events) {
    LIDAR_REQUIRE(caps != nullptr, LIDAR_ERR_INVALID, "lidar_frame_voxel_density: caps is NULL");
    LIDAR_REQUIRE(n >= 0 && n <= caps->max_points, LIDAR_ERR_CAPACITY,
                  "lidar_frame_voxel_density: n=%lld exceeds caps.max_points=%lld", (long long)n,
                  (long long)caps->max_points);
    LIDAR_REQUIRE(n < (1ll << 31), LIDAR_ERR_CAPACITY, "lidar_frame_voxel_density: n must be < 2^31");
    LIDAR_REQUIRE(voxel_size > 0.0 && voxel_size == voxel_size, LIDAR_ERR_INVALID,
                  "lidar_frame_voxel_density: voxel_size must be > 0");
    LIDAR_REQUIRE(grid_size >= 0.0, LIDAR_ERR_INVALID, "lidar_frame_voxel_density: grid_size must be >= 0");
    LIDAR_REQUIRE(caps->max_key_space > 0 && caps->max_key_space < (1ll << 31), LIDAR_ERR_INVALID,
                  "lidar_frame_voxel_density: caps.max_key_space must be in (0, 2^31)");
    LIDAR_REQUIRE(d_desc && d_voxel_key && d_inverse && d_voxels, LIDAR_ERR_INVALID,
                  "lidar_frame_voxel_density: NULL output");
    LIDAR_REQUIRE(grid_size == 0.0 || (d_grid && caps->max_nx > 0 && caps->max_ny > 0), LIDAR_ERR_INVALID,
                  "lidar_frame_voxel_density: density grid requested without d_grid / capacities");
    LIDAR_REQUIRE(n == 0 || d_points, LIDAR_ERR_INVALID, "lidar_frame_voxel_density: NULL points");
    const FrameWsLayout L = frame_layout(*caps);
    LIDAR_REQUIRE(d_ws && ws_bytes >= L.total, LIDAR_ERR_WORKSPACE,
                  "lidar_frame_voxel_density: workspace too small (%zu < %zu)", ws_bytes, L.total);
    LIDAR_REQUIRE(L.tiles <= 65536, LIDAR_ERR_CAPACITY, "lidar_frame_voxel_density: caps.max_key_space too large");
    char* ws = static_cast<char*>(d_ws);
    double* partial = reinterpret_cast<double*>(ws + L.off_partial);
    FrameCtrl* ctrl = reinterpret_cast<FrameCtrl*>(ws + L.off_ctrl);
    uint32_t* groups = reinterpret_cast<uint32_t*>(ws + L.off_groups);
    unsigned long long* tile_desc = reinterpret_cast<unsigned long long*>(ws + L.off_tile_desc);
    long long* acc = reinterpret_cast<long long*>(ws + L.off_acc);
    int32_t* cnt = reinterpret_cast<int32_t*>(ws + L.off_cnt);
    uint32_t* filter = reinterpret_cast<uint32_t*>(ws + L.off_filter);

    FrameParams P;
    P.pts = static_cast<const float4*>(d_points);
    P.n = n;
    P.voxel = voxel_size;
    P.grid = grid_size;
    P.has_origin = h_origin3 != nullptr;
    P.has_range = h_xy_range4 != nullptr;
    for (int c = 0; c < 3; ++c) P.origin[c] = h_origin3 ? h_origin3[c] : 0.0;
    for (int c = 0; c < 4; ++c) P.xyr[c] = h_xy_range4 ? h_xy_range4[c] : 0.0;
    P.max_key_space = caps->max_key_space;
    P.max_nx = caps->max_nx;
    P.max_ny = caps->max_ny;
    int lg = 0;
    while ((1ll << lg) < (n > 1 ? n : 1)) ++lg;
    P.fix_bits_budget = 62 - lg;

    cudaStream_t st = as_stream(stream);
    auto mark = [&](int i) -> cudaError_t {
        return events ? cudaEventRecord(static_cast<cudaEvent_t>(events[i]), st) : cudaSuccess;
    };
    const int grid_cap = grid_size > 0.0 ? caps->max_nx * caps->max_ny : 0;
    LIDAR_CUDA_TRY(mark(0));
    int bgrid = frame_grid(n > 0 ? n : 1, 4);
    if (bgrid < sm_count()) bgrid = sm_count();   // enough CTAs to zero the bitmap quickly
    if (bgrid > kBboxMaxBlocks) bgrid = kBboxMaxBlocks;
    k_frame_prep<<<bgrid, kFrameThreads, 0, st>>>(P, partial, ctrl, d_desc, d_grid, grid_cap, tile_desc, L.tiles,
                                                  groups, L.groups, filter);
    LIDAR_CHECK_LAUNCH();
    LIDAR_CUDA_TRY(mark(1));
    if (n == 0) {
        for (int i = 2; i <= 5; ++i) LIDAR_CUDA_TRY(mark(i));
        return LIDAR_OK;
    }
    k_frame_mark<<<frame_grid(n, 2), kFrameThreads, 0, st>>>(P.pts, d_desc, d_voxel_key, groups, filter, d_grid);
    LIDAR_CHECK_LAUNCH();
    LIDAR_CUDA_TRY(mark(2));
    {
        // one CTA per tile when they all fit in a single wave, so nobody spins on an unscheduled tile
        int sgrid = sm_count() * 8;
        if ((int64_t)sgrid > L.tiles) sgrid = (int)L.tiles;
        k_frame_scan<<<sgrid, kFrameThreads, 0, st>>>(groups, tile_desc, ctrl, d_desc);
        LIDAR_CHECK_LAUNCH();
    }
    LIDAR_CUDA_TRY(mark(3));
    k_frame_rank<<<frame_grid(n, 2), kFrameThreads, 0, st>>>(P.pts, d_desc, d_voxel_key, groups, filter, d_inverse,
                                                              acc, cnt, d_voxels);
    LIDAR_CHECK_LAUNCH();
    LIDAR_CUDA_TRY(mark(4));
    k_frame_finalize<<<frame_grid(n, 4), kFrameThreads, 0, st>>>(d_desc, acc, cnt, d_voxels);
    LIDAR_CHECK_LAUNCH();
    LIDAR_CUDA_TRY(mark(5));
    return LIDAR_OK;
}

int lidar_frame_voxel_density(const void* d_points, int64_t n, double voxel_size, double grid_size,
                              const double* h_origin3, const double* h_xy_range4, int32_t* d_voxel_key,
                              int32_t* d_inverse, lidar_voxel* d_voxels, int32_t* d_grid, lidar_frame_desc* d_desc,
                              const lidar_frame_caps* caps, void* d_ws, size_t ws_bytes, void* stream) {
    return frame_voxel_density_impl(d_points, n, voxel_size, grid_size, h_origin3, h_xy_range4, d_voxel_key,
                                    d_inverse, d_voxels, d_grid, d_desc, caps, d_ws, ws_bytes, stream, nullptr);
}

int lidar_frame_voxel_density_timed(const void* d_points, int64_t n, double voxel_size, double grid_size,
                                    const double* h_origin3, const double* h_xy_range4, int32_t* d_voxel_key,
                                    int32_t* d_inverse, lidar_voxel* d_voxels, int32_t* d_grid,
                                    lidar_frame_desc* d_desc, const lidar_frame_caps* caps, void* d_ws,
                                    size_t ws_bytes, void* stream, void** h_events6) {
    LIDAR_REQUIRE(h_events6 != nullptr, LIDAR_ERR_INVALID, "lidar_frame_voxel_density_timed: events is NULL");
    return frame_voxel_density_impl(d_points, n, voxel_size, grid_size, h_origin3, h_xy_range4, d_voxel_key,
                                    d_inverse, d_voxels, d_grid, d_desc, caps, d_ws, ws_bytes, stream, h_events6);
}

}  // extern "C"
