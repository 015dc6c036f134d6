// K5 (+K6 fused) — one LiDAR frame: bbox -> voxel downsample (+ ground-plane density histogram).
//
// Voxel downsample is a NEW op (absent from the reference; contract = SURVEY.md Appendix B.1); the
// fused density grid has calculate_grid_density semantics (utils/data_processing.py:282-328):
// margin 2g, np.arange edges (fill rule e(i) = a + i*fl(fl(a+g)-a), Appendix A.2), histogramdd
// binning (Appendix A.1), counts indexed [x][y].
//
// Design: "occupancy-bitmap ranking" instead of a key sort.
//   The voxel key space (Dx*Dy*Dz cells) is held as a bitmap (1 bit per cell, 22 MB for a
//   100 m x 100 m x 2 m frame at 0.05 m — L2 resident on B200).  The rank of a voxel in ascending
//   key order is the number of occupied cells before it, i.e. a popcount prefix over the bitmap.
//   That yields exactly the output order of a stable sort-by-key + segmented reduce, without
//   moving a single point, and every step is order independent:
//     prep      min/max of x,y,z,intensity; the frame descriptor (origin, dims, key space, fixed-point
//               scales, histogram edges) is derived ON DEVICE
//     mark      per point: voxel key, atomicOr into the bitmap, density bin -> RED.ADD into the grid.
//               The atomicOr returns the old word: a point that finds its bit already set is a LATER
//               member of its voxel ("dup"); the flag travels with the key (bit 31).
//     scan      scan of bitmap popcounts -> exclusive prefix stored inside each 32-byte group
//     rank      per point: rank = prefix + popc(bits below) -> inverse.  The FIRST member of a voxel
//               (94 % of the points of a crowd frame are the only member) stores itself as the voxel
//               record with one 32-byte store; dup members accumulate (p - ref) in 2^-k fixed point
//               with integer atomics.
//     finalize  voxels with dup members only: centroid = ref + (first + sum)/count; re-zeroes accumulators
//   Integer accumulation makes centroids independent of the order in which atomics land:
//   bit-identical run to run, and (p - ref)*2^k is an exact integer, so the sums are exact.
//
// Two back ends with identical arithmetic (shared device functions) and identical outputs:
//   k_frame_fused   ONE persistent cooperative kernel: the frame is pulled into shared memory once by
//                   the TMA bulk-copy engine and all phases run on the resident points, separated by
//                   grid barriers (default)
//   k_frame_prep/mark/scan/rank/finalize   five dependent kernels (any frame size, no co-residency)
//
// Nothing in a frame needs the host: capacities are fixed per stream of frames, the descriptor is
// read back together with the results.
#include "common.cuh"
#include "edges.cuh"

namespace lidar {

constexpr int kFrameThreads = 256;
// Occupancy groups: one 32-byte sector = {prefix, 7 x 32 occupancy bits} = 224 voxels.  Packing the
// popcount prefix next to the bits means the rank of a point costs ONE random sector read.
constexpr int kGroupVoxels = 224;
constexpr int kScanGroupsPerThread = 4;                               // each thread scans 4 adjacent groups
constexpr int kScanTileGroups = kFrameThreads * kScanGroupsPerThread; // 1024 groups = 229 376 voxels per tile

struct FrameWsLayout {
    size_t off_partial, off_ctrl, off_groups, off_tile_desc, off_acc, off_cnt, off_cta_desc, off_trace, off_grid_rep, off_l1, off_l1cnt,
           off_ptotal, off_pvox, off_bucket, off_prank, total;
    int64_t groups, tiles, l1_words;
};

struct FrameCtrl {
    unsigned int bbox_ticket;
    unsigned int scan_ticket;
    unsigned int grid_bar;     // fused kernel: grid-barrier arrival counter (reset by the last CTA to exit)
    unsigned int grid_bar_b;   // scan-order variant: the extra barrier of the sparse scan (same reset)
    unsigned int exit_ticket;  // fused kernel: CTAs that have passed the last barrier
    unsigned long long dirty_groups;  // occupancy groups the previous frame may have set
};

constexpr int kFusedMaxCtas = 1024;   // cap of the fused kernel's grid (one slot of cta_desc each)
// The density REDs of a 1 M-point frame land in a 166 KB grid: same-sector atomics serialise in L2
// (micro-benchmark: 16.4 us for 1 M REDs into one grid, 11.2 us into 8 replicas).  The fused kernel
// spreads them over up to 8 private replicas (CTA b uses replica b % R) and merges them after the mark
// phase.  The replica area is all-zero between frames.
constexpr int kGridRepCells = 1 << 19;   // 2 MB
constexpr int kGridRepMax = 8;

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
    L.off_cta_desc = take(sizeof(unsigned long long) * kFusedMaxCtas);
    L.off_trace = take(sizeof(unsigned long long) * 16 * kFusedMaxCtas);
    L.off_grid_rep = take(sizeof(int32_t) * kGridRepCells);
    L.l1_words = (L.groups + 31) / 32;             // summary: one bit per occupancy group (scan-order variant)
    L.off_l1 = take(sizeof(uint32_t) * (size_t)L.l1_words);
    L.off_l1cnt = take(sizeof(uint32_t) * (size_t)L.l1_words);   // occupied cells per summary word, then their prefix
    // partitioned back end (k_frame_part): claim counters and voxel counts per partition, the bucket of {key, cell}
    // entries grouped by partition and the owners' answers (2048 = kPartMax partitions of 2^18 cells)
    L.off_ptotal = take(sizeof(uint32_t) * 2048);
    L.off_pvox = take(sizeof(uint32_t) * 2048);
    L.off_bucket = take(sizeof(uint2) * (size_t)c.max_points);
    L.off_prank = take(sizeof(uint32_t) * (size_t)c.max_points);
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

// ---- descriptor derivation --------------------------------------------------------------------
// Called by ALL threads of a CTA (>= 128 threads) with the frame's bounding box in shared memory: the
// independent pieces (three voxel axes, two histogram axes, fixed-point scales, two magic divisors) run
// on different warps instead of one after the other on one thread (~3 us -> ~1 us on the critical path
// of every frame).  `s_stat` is 8 ints of shared scratch.  Ends with a __syncthreads().
__device__ __forceinline__ int first_error(int a, int b) { return a ? a : b; }

__device__ void derive_desc_cta(const FrameParams& P, const double* bb, lidar_frame_desc* D, int* s_stat) {
    const int tid = threadIdx.x;
    const int part = ((tid & 31) < 2 && (tid >> 5) < 4) ? (tid >> 5) + 4 * (tid & 31) : -1;   // 0..7
    if (part >= 0) s_stat[part] = 0;
    if (part >= 0 && part < 3) {
        // voxel axis c: origin, extent in voxels
        const int c = part;
        int status = 0;
        const double o = P.has_origin ? P.origin[c] : bb[c];
        D->origin[c] = o;
        if (bb[c] < o) status = LIDAR_ERR_INVALID;  // a point below the origin would index < 0
        const double span = floor(__ddiv_rn(__dsub_rn(bb[4 + c], o), P.voxel));
        long long d = (span >= 0.0 && span < 2147483000.0) ? (long long)span + 1 : 0;
        if (d <= 0) { d = 1; if (P.n > 0) status = first_error(status, LIDAR_ERR_CAPACITY); }
        D->dims[c] = (int)d;
        s_stat[part] = status;
    } else if (part == 3) {
        for (int c = 0; c < 4; ++c) {
            D->bbox_min[c] = bb[c];
            D->bbox_max[c] = bb[4 + c];
        }
        D->voxel = P.voxel;
        D->n_points = P.n;
        D->n_voxels = 0;
        for (int k = 0; k < 16; ++k) D->trace_ns[k] = 0u;
        D->origin[3] = 0.0;
        D->dims[3] = 0;
        D->grid = P.grid;
        // fixed point: |p - ref| < 2*voxel + 2^-24, n members at most  ->  that * 2^k * n < 2^62
        int e;
        frexp(P.voxel * 2.0 + 0x1p-24, &e);  // bound < 2^e
        D->fix_scale_xyz = ldexp(1.0, P.fix_bits_budget - e);
        double wmax = fmax(fabs(bb[3]), fabs(bb[7]));
        if (!(wmax > 0.0) || !isfinite(wmax)) wmax = 1.0;
        frexp(wmax, &e);
        D->fix_scale_w = ldexp(1.0, P.fix_bits_budget - e - 1);
    } else if (part == 4 || part == 5) {
        // density grid edges of axis c, numpy arange rule (data_processing.py:305-313)
        const int c = part - 4;
        int status = 0;
        double a = 0.0, e1 = 0.0, delta = 0.0;
        int nb = 0;
        if (P.grid > 0.0) {
            const double lo = P.has_range ? P.xyr[2 * c] : bb[c];
            const double hi = P.has_range ? P.xyr[2 * c + 1] : bb[4 + c];
            const ArangeAxis ax = arange_axis(lo, hi, P.grid, c == 0 ? P.max_nx : P.max_ny, P.n > 0);
            a = ax.a; e1 = ax.e1; delta = ax.delta; nb = ax.nb; status = ax.status;
        }
        if (c == 0) { D->ex0 = a; D->ex1 = e1; D->exd = delta; D->nx = nb; }
        else        { D->ey0 = a; D->ey1 = e1; D->eyd = delta; D->ny = nb; }
        s_stat[part] = status;
    }
    __syncthreads();
    if (part == 0) {
        // key space (overflow-safe product) and the combined status, first error in axis order wins
        int status = first_error(first_error(s_stat[0], s_stat[1]), s_stat[2]);
        long long ks = 1;
        for (int c = 0; c < 3; ++c) {
            if (ks > (1ll << 40)) status = first_error(status, LIDAR_ERR_CAPACITY);
            ks *= (long long)D->dims[c];
        }
        D->key_space = ks;
        if (ks > P.max_key_space || ks >= (1ll << 31)) status = first_error(status, LIDAR_ERR_CAPACITY);
        status = first_error(status, first_error(s_stat[4], s_stat[5]));
        if (P.n == 0) status = 0;
        D->status = status;
        // the fp32 index guess needs an origin that fp32 represents exactly (true for a bbox-derived origin)
        int fast = 1;
        for (int c = 0; c < 3; ++c)
            if ((double)(float)D->origin[c] != D->origin[c]) fast = 0;
        if ((double)(float)P.voxel == 0.0) fast = 0;
        D->fast_f32 = fast;
    } else if (part == 1 || part == 2) {
        // Granlund-Montgomery: for 0 <= n < 2^31 and 2^(l-1) < d <= 2^l,  n / d == (n * ceil(2^(31+l)/d)) >> (31+l)
        const int c = part - 1;
        const unsigned long long d = (unsigned long long)(c == 0 ? D->dims[2] : D->dims[1]);
        int l = 0;
        while ((1ull << l) < d) ++l;
        const unsigned long long num = 1ull << (31 + l);
        const unsigned m = (unsigned)((num + d - 1) / d);
        if (c == 0) { D->magic_dz = m; D->shift_dz = l; } else { D->magic_dy = m; D->shift_dy = l; }
    }
    __syncthreads();
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
             int64_t groups_cap) {
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
    __shared__ int s_stat[8];
    derive_desc_cta(P, s_bb, D, s_stat);
    if (threadIdx.x == 0) {
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

// fire-and-forget reductions: spelled in PTX so that ptxas emits REDG (no return path) and never
// ATOMG with a discarded result
__device__ __forceinline__ void red_add_s32(int32_t* p, int v) {
    asm volatile("red.relaxed.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_add_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// ---- per-point arithmetic shared by both back ends -----------------------------------------------
struct MarkConst {
    float of[3], rvf, axf, ayf, rdxf, rdyf;
    double rv, rdx, rdy;
    int fast, do_grid;
};

__device__ __forceinline__ MarkConst make_mark_const(const lidar_frame_desc& D) {
    MarkConst K;
    K.do_grid = D.grid > 0.0;
    K.rv = __ddiv_rn(1.0, D.voxel);
    K.rdx = K.do_grid ? __ddiv_rn(1.0, D.exd) : 0.0;
    K.rdy = K.do_grid ? __ddiv_rn(1.0, D.eyd) : 0.0;
    K.of[0] = (float)D.origin[0]; K.of[1] = (float)D.origin[1]; K.of[2] = (float)D.origin[2];
    K.rvf = (float)K.rv;
    K.axf = (float)D.ex0; K.ayf = (float)D.ey0; K.rdxf = (float)K.rdx; K.rdyf = (float)K.rdy;
    K.fast = D.fast_f32;
    return K;
}

// B.1: key = (ix*Dy + iy)*Dz + iz with i = floor((f64(p) - origin) / voxel)
__device__ __forceinline__ int voxel_key_of(const float4& q, const lidar_frame_desc& D, const MarkConst& K) {
    int ix, iy, iz;
    const bool fx = K.fast && fast_voxel_index(q.x, K.of[0], K.rvf, ix);
    const bool fy = K.fast && fast_voxel_index(q.y, K.of[1], K.rvf, iy);
    const bool fz = K.fast && fast_voxel_index(q.z, K.of[2], K.rvf, iz);
    if (!fx) ix = floor_div_exact(__dsub_rn((double)q.x, D.origin[0]), D.voxel, K.rv);
    if (!fy) iy = floor_div_exact(__dsub_rn((double)q.y, D.origin[1]), D.voxel, K.rv);
    if (!fz) iz = floor_div_exact(__dsub_rn((double)q.z, D.origin[2]), D.voxel, K.rv);
    return (ix * D.dims[1] + iy) * D.dims[2] + iz;
}
// calculate_grid_density cell of a point ([x][y] layout), -1 when outside the edges
__device__ __forceinline__ int grid_cell_of(const float4& q, const lidar_frame_desc& D, const MarkConst& K) {
    const int bx = fast_arange_bin(q.x, (double)q.x, K.axf, K.rdxf, D.ex0, D.ex1, D.exd, K.rdx, D.nx);
    const int by = fast_arange_bin(q.y, (double)q.y, K.ayf, K.rdyf, D.ey0, D.ey1, D.eyd, K.rdy, D.ny);
    return (bx >= 0 && by >= 0) ? bx * D.ny + by : -1;
}

constexpr unsigned kDupFlag = 0x80000000u;   // keys are < 2^31: bit 31 carries "a later member of its voxel"

// occupancy bit of a key: word address inside its group and the bit mask
__device__ __forceinline__ uint32_t* occupancy_word(uint32_t* groups, int key, unsigned& bit) {
    const unsigned g = (unsigned)key / kGroupVoxels, b = (unsigned)key - g * kGroupVoxels;
    bit = 1u << (b & 31);
    return groups + (size_t)g * 8 + 1 + (b >> 5);
}
__device__ __forceinline__ void red_or_b32(unsigned* p, unsigned v) {
    asm volatile("red.relaxed.gpu.global.or.b32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// 256-bit global accesses (sm_100: LDG.E.ENL2.256 / STG.E.ENL2.256): a whole 32-byte occupancy group or
// voxel record moves with ONE request, i.e. one L1 wavefront per lane instead of two
struct __align__(32) Word8 { unsigned w[8]; };
__device__ __forceinline__ Word8 ld_cg_256(const void* p) {
    Word8 v;
    asm volatile("ld.global.cg.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v.w[0]), "=r"(v.w[1]), "=r"(v.w[2]), "=r"(v.w[3]), "=r"(v.w[4]), "=r"(v.w[5]), "=r"(v.w[6]), "=r"(v.w[7])
                 : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_256(void* p, const Word8& v) {
    asm volatile("st.global.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"l"(p), "r"(v.w[0]), "r"(v.w[1]), "r"(v.w[2]), "r"(v.w[3]), "r"(v.w[4]), "r"(v.w[5]), "r"(v.w[6]), "r"(v.w[7])
                 : "memory");
}
__device__ __forceinline__ unsigned group_popc(const Word8& g) {
    return __popc(g.w[1]) + __popc(g.w[2]) + __popc(g.w[3]) + __popc(g.w[4]) + __popc(g.w[5]) + __popc(g.w[6]) + __popc(g.w[7]);
}
// rank of a key = prefix of its group + occupied cells below it inside the group
__device__ __forceinline__ unsigned rank_in_group(const Word8& g, unsigned b) {
    const int wi = 1 + (int)(b >> 5);
    unsigned r = g.w[0];
#pragma unroll
    for (int k = 1; k < 8; ++k) {
        if (k < wi) r += __popc(g.w[k]);
        else if (k == wi) r += __popc(g.w[k] & ((1u << (b & 31)) - 1u));
    }
    return r;
}
__device__ __forceinline__ unsigned rank_of(int key, const uint32_t* __restrict__ groups) {
    const unsigned g = (unsigned)key / kGroupVoxels, b = (unsigned)key - g * kGroupVoxels;
    return rank_in_group(ld_cg_256(groups + (size_t)g * 8), b);
}
// the first member of a voxel stores itself as the record (count 1): one full-sector store
__device__ __forceinline__ void store_first_member(lidar_voxel* __restrict__ voxels, unsigned r, const float4& q, int key) {
    Word8 rec;
    rec.w[0] = __float_as_uint(q.x); rec.w[1] = __float_as_uint(q.y);
    rec.w[2] = __float_as_uint(q.z); rec.w[3] = __float_as_uint(q.w);
    rec.w[4] = 1u; rec.w[5] = (unsigned)key; rec.w[6] = 0u; rec.w[7] = 0u;
    st_256(voxels + r, rec);
}

struct Fixed4 { long long x, y, z, w; };
// (p - ref) * 2^k of a member of voxel `key`: an exact integer (see voxel_ref)
__device__ __forceinline__ Fixed4 fixed_offsets(const float4& q, int key, const lidar_frame_desc& D) {
    int ix, iy, iz;
    decode_key(key, D, ix, iy, iz);
    const double cx = voxel_ref(D.origin[0], ix, D.voxel);
    const double cy = voxel_ref(D.origin[1], iy, D.voxel);
    const double cz = voxel_ref(D.origin[2], iz, D.voxel);
    Fixed4 f;
    f.x = __double2ll_rn(__dmul_rn(__dsub_rn((double)q.x, cx), D.fix_scale_xyz));
    f.y = __double2ll_rn(__dmul_rn(__dsub_rn((double)q.y, cy), D.fix_scale_xyz));
    f.z = __double2ll_rn(__dmul_rn(__dsub_rn((double)q.z, cz), D.fix_scale_xyz));
    f.w = __double2ll_rn(__dmul_rn((double)q.w, D.fix_scale_w));
    return f;
}
// a later ("dup") member adds itself to the accumulators of its voxel; results unused => RED
__device__ __forceinline__ void accumulate_dup(const float4& q, int key, unsigned r, const lidar_frame_desc& D,
                                               long long* __restrict__ acc, int32_t* __restrict__ cnt) {
    const Fixed4 f = fixed_offsets(q, key, D);
    unsigned long long* A = reinterpret_cast<unsigned long long*>(acc + (size_t)r * 4);
    red_add_u64(A + 0, (unsigned long long)f.x);
    red_add_u64(A + 1, (unsigned long long)f.y);
    red_add_u64(A + 2, (unsigned long long)f.z);
    red_add_u64(A + 3, (unsigned long long)f.w);
    red_add_s32(cnt + r, 1);
}
// voxel r has cnt[r] dup members plus the first member sitting in its record
__device__ __forceinline__ void finalize_voxel(unsigned r, const lidar_frame_desc& D, double isx, double isw,
                                               long long* __restrict__ acc, int32_t* __restrict__ cnt,
                                               lidar_voxel* __restrict__ voxels) {
    const int c = __ldcg(cnt + r) + 1;
    Word8 rec = ld_cg_256(voxels + r);
    const float4 first = make_float4(__uint_as_float(rec.w[0]), __uint_as_float(rec.w[1]), __uint_as_float(rec.w[2]),
                                     __uint_as_float(rec.w[3]));
    const int key = (int)rec.w[5];
    const Fixed4 f = fixed_offsets(first, key, D);
    longlong2* A = reinterpret_cast<longlong2*>(acc + (size_t)r * 4);
    const longlong2 s01 = __ldcg(A), s23 = __ldcg(A + 1);
    int ix, iy, iz;
    decode_key(key, D, ix, iy, iz);
    const double dc = (double)c;
    const double cx = voxel_ref(D.origin[0], ix, D.voxel);
    const double cy = voxel_ref(D.origin[1], iy, D.voxel);
    const double cz = voxel_ref(D.origin[2], iz, D.voxel);
    rec.w[0] = __float_as_uint((float)__dadd_rn(cx, __ddiv_rn(__dmul_rn((double)(s01.x + f.x), isx), dc)));
    rec.w[1] = __float_as_uint((float)__dadd_rn(cy, __ddiv_rn(__dmul_rn((double)(s01.y + f.y), isx), dc)));
    rec.w[2] = __float_as_uint((float)__dadd_rn(cz, __ddiv_rn(__dmul_rn((double)(s23.x + f.z), isx), dc)));
    rec.w[3] = __float_as_uint((float)__ddiv_rn(__dmul_rn((double)(s23.y + f.w), isw), dc));
    rec.w[4] = (unsigned)c;
    st_256(voxels + r, rec);
    // restore the all-zero invariant of the accumulators for the next frame
    A[0] = make_longlong2(0, 0);
    A[1] = make_longlong2(0, 0);
    cnt[r] = 0;
}

// ---- k_frame_mark -----------------------------------------------------------------------------
__global__ void __launch_bounds__(kFrameThreads)
k_frame_mark(const float4* __restrict__ pts, const lidar_frame_desc* __restrict__ Dg,
             int32_t* __restrict__ voxel_key, uint32_t* __restrict__ groups, int32_t* __restrict__ grid_out) {
    __shared__ lidar_frame_desc D;
    LoadF32x4 L{pts};
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (threadIdx.x == 0) D = *Dg;
    __syncthreads();
    if (D.status != 0) return;
    const int64_t n = D.n_points;
    const MarkConst K = make_mark_const(D);
    // two points per trip: both atomicOr are in flight before either result is consumed
    for (int64_t i = t0; i < n; i += 2 * stride) {
        const bool two = i + stride < n;
        const float4 a = L.raw(i);
        const float4 b = two ? L.raw(i + stride) : a;
        const int ka = voxel_key_of(a, D, K), kb = voxel_key_of(b, D, K);
        unsigned bita, bitb, olda, oldb = 0u;
        uint32_t* wa = occupancy_word(groups, ka, bita);
        uint32_t* wb = occupancy_word(groups, kb, bitb);
        olda = atomicOr(wa, bita);
        if (two) oldb = atomicOr(wb, bitb);
        if (K.do_grid) {
            const int ca = grid_cell_of(a, D, K);
            if (ca >= 0) red_add_s32(grid_out + ca, 1);
            if (two) {
                const int cb = grid_cell_of(b, D, K);
                if (cb >= 0) red_add_s32(grid_out + cb, 1);
            }
        }
        // the dup flag rides on the key; k_frame_rank strips it again
        st_stream_s32(voxel_key + i, (int)((unsigned)ka | ((olda & bita) ? kDupFlag : 0u)));
        if (two) st_stream_s32(voxel_key + i + stride, (int)((unsigned)kb | ((oldb & bitb) ? kDupFlag : 0u)));
    }
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
// The accumulate path is needed by ~6 % of the points but, taken in place, would be executed by almost
// every warp (divergence).  Each warp therefore parks those points in a private shared-memory ring and
// drains it 32 at a time with all lanes active.
struct DupItem {
    float4 q;
    int key;
    unsigned r;
    int pad[2];
};
constexpr int kRingSize = 64;

__global__ void __launch_bounds__(kFrameThreads)
k_frame_rank(const float4* __restrict__ pts, const lidar_frame_desc* __restrict__ Dg,
             int32_t* __restrict__ voxel_key, const uint32_t* __restrict__ groups,
             int32_t* __restrict__ inverse, long long* __restrict__ acc,
             int32_t* __restrict__ cnt, lidar_voxel* __restrict__ voxels) {
    __shared__ lidar_frame_desc D;
    __shared__ DupItem s_ring[kFrameThreads / 32][kRingSize];
    if (threadIdx.x == 0) D = *Dg;
    __syncthreads();
    if (D.status != 0) return;
    const int64_t n = D.n_points;
    LoadF32x4 L{pts};
    const unsigned lane = lane_id();
    DupItem* ring = s_ring[threadIdx.x >> 5];
    unsigned head = 0, count = 0;     // warp-uniform
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t n_round = ((n + 31) / 32) * 32;   // keep whole warps in the loop for the ballots
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
        const bool live = i < n;
        bool dup = false;
        DupItem it;
        if (live) {
            const unsigned fk = (unsigned)__ldcg(voxel_key + i);   // written by k_frame_mark, flag in bit 31
            dup = (fk & kDupFlag) != 0u;
            it.key = (int)(fk & ~kDupFlag);
            it.q = L.raw(i);
            it.r = rank_of(it.key, groups);
            st_stream_s32(inverse + i, (int)it.r);
            if (dup) voxel_key[i] = it.key;                          // strip the flag (6 % of the points)
            else store_first_member(voxels, it.r, it.q, it.key);
        }
        const unsigned m = __ballot_sync(0xffffffffu, dup);
        if (m) {
            if (dup) ring[(head + count + __popc(m & lanemask_lt())) % kRingSize] = it;
            count += __popc(m);
            __syncwarp();
            if (count >= 32) {
                const DupItem& e = ring[(head + lane) % kRingSize];
                accumulate_dup(e.q, e.key, e.r, D, acc, cnt);
                head = (head + 32) % kRingSize;
                count -= 32;
                __syncwarp();
            }
        }
    }
    if (lane < count) {
        const DupItem& e = ring[(head + lane) % kRingSize];
        accumulate_dup(e.q, e.key, e.r, D, acc, cnt);
    }
}

// ---- k_frame_finalize -------------------------------------------------------------------------
// cnt[r] != 0 exactly for the voxels with more than one member: a coalesced sweep over cnt finds them;
// the ~6 % of hits are parked in a per-warp ring and finished 32 at a time, so the fp64 divisions run
// with full warps instead of being executed, mostly masked, by every warp.
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
        const bool hit = r < V && __ldcg(cnt + r) != 0;
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

// ================================================================================================
// k_frame_fused — the whole frame in ONE persistent cooperative kernel, points resident on chip.
//
// One CTA per SM (grid = a multiple of the SM count, all co-resident: cooperative launch).  Every CTA
// owns a contiguous chunk of the frame and pulls it into shared memory ONCE with the TMA bulk-copy
// engine (cp.async.bulk + mbarrier, SASS UBLKCP): 1 M points / 148 SMs = 6 757 points = 106 KB per SM —
// a whole frame fits in the 33 MB of shared memory of the chip.  All later phases read the points (and
// their voxel keys) from shared memory, so HBM sees the compulsory traffic only: 16 B/point in,
// key + inverse + voxel records out.  Points beyond the shared-memory capacity of a CTA ("spill") are
// re-read from global memory (L2) in each phase, so any frame size works.
//
//   phase 0  TMA load of the chunk; zero the density grid; min/max of the chunk -> partials
//   -- grid barrier --   every CTA folds the partials and derives the frame descriptor redundantly
//   phase 1  mark: voxel key (+dup flag) -> smem, key -> global, atomicOr occupancy bit, density RED;
//            four points per thread in flight
//   -- grid barrier --
//   phase 2  scan: each CTA popcounts a contiguous range of occupancy groups (one 256-bit load per
//            group, warps walk contiguous rows), publishes its total (flag + value), waits for the
//            totals of ALL CTAs (co-resident => spinning is safe), writes the exclusive prefix into
//            word 0 of each of its groups
//   -- grid barrier --
//   phase 3  rank: prefix + popc(bits below) -> inverse; first members store their record, dup
//            members accumulate exact fixed-point sums
//   -- grid barrier --
//   phase 4  re-zero this CTA's range of the occupancy bitmap (clean for the next frame) and finalize
//            the multi-member voxels (sweep of the member counters)
//
// Same arithmetic (device functions) as the five-kernel path => identical outputs.
// ================================================================================================
constexpr int kFusedMaxThreads = 512;     // 128 registers per thread: room for four points in flight
constexpr int kFusedLoadStages = 4;       // the chunk arrives in 4 bulk copies, each with its own mbarrier
constexpr int kFusedRing = 64;
constexpr int kFusedBatch = 4;            // points per thread per trip in the mark and rank phases

struct FusedArgs {
    FrameParams P;
    double* partial;
    FrameCtrl* ctrl;
    lidar_frame_desc* D;
    int32_t* voxel_key;
    int32_t* inverse;
    lidar_voxel* voxels;
    int32_t* grid_out;
    int grid_cap;
    uint32_t* groups;
    int64_t groups_cap;
    long long* acc;
    int32_t* cnt;
    unsigned long long* cta_desc;
    unsigned long long* trace_all;   // [G][16] %globaltimer stamps of every CTA (diagnostics)
    int32_t* grid_rep;               // kGridRepCells zeroed cells: private replicas of the density grid
    uint32_t* l1;                    // scan-order variant: summary bitmap, bit g = occupancy group g is not empty
    uint32_t* l1cnt;                 // scan-order variant: occupied cells per summary word, then their exclusive prefix
    int early_load;       // programmatic dependent launch: 1 = the frame was complete in memory before the launch was
                          // enqueued, so the TMA load and the bounding box may run BEFORE griddepcontrol.wait
    int smem_points;      // resident points per CTA (shared-memory capacity)
    int smem_groups;      // capacity of the per-group popcount cache (bytes) per CTA
};

__device__ __forceinline__ unsigned fused_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void fused_mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(fused_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fused_mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fused_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fused_mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(fused_smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void fused_bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(fused_smem_u32(dst)), "l"(src), "r"(bytes), "r"(fused_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_add_u32(unsigned* p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// a spin that lasts two seconds means the grid is not co-resident (or a CTA died): kill the context
// instead of hanging the device
struct SpinGuard {
    unsigned long long t0;
    unsigned polls;
    __device__ __forceinline__ SpinGuard() : t0(0ull), polls(0u) {}
    __device__ __forceinline__ void tick() {
        if ((++polls & 0x3ffu) == 0u) {
            const unsigned long long t = global_timer_ns();
            if (t0 == 0ull) t0 = t;
            else if (t - t0 > 2000000000ull) __trap();
        }
    }
};

// All CTAs of the grid are co-resident (cooperative launch), so a counter barrier cannot deadlock.
// `target` = arrivals expected so far = barrier ordinal * gridDim.x; the counter is reset to zero by
// the last CTA to leave the kernel.
__device__ __forceinline__ void fused_grid_barrier(unsigned* ctr, unsigned target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        red_release_add_u32(ctr, 1u);
        SpinGuard guard;
        while (ld_acquire_u32(ctr) < target) guard.tick();
    }
    __syncthreads();
}

#define FUSED_TRACE(k) do { if (tid == 0) s_trace[k] = global_timer_ns(); } while (0)

// kScan = the scan-order variant (lidar_frame_set_fused_scan_order): frames as a sensor delivers them, and / or key
// spaces much larger than the data.  It adds (a) run-length aggregation across adjacent lanes in the mark and rank
// phases and (b) a scan / clean that walks a summary bitmap of the occupied groups instead of streaming every group.
// The default instantiation (kScan = false) contains none of that code: the kernel is instruction-cache bound, and
// compiling both into one body cost the shuffled benchmark 5 %.
// kMaxT = the largest block the instantiation may be launched with: 512 (up to 128 registers per thread: four points in
// flight without spills) or 1024 (64 registers, ~100 bytes of spills, but 32 resident warps per SM instead of 16 -- the
// kernel is bound by dependent-instruction latency, profiles/r2_ncu_fused_streaming_raw.csv).
template <bool kScan, int kMaxT>
__global__ void __launch_bounds__(kMaxT, 1)
k_frame_fused(const FusedArgs A) {
    extern __shared__ __align__(128) unsigned char fsm[];
    __shared__ lidar_frame_desc D;
    __shared__ __align__(8) unsigned long long s_bar[kFusedLoadStages];
    __shared__ double s_red[kMaxT / 32][8];
    __shared__ double s_bb[8];
    __shared__ unsigned s_wsum[kMaxT / 32];
    __shared__ unsigned long long s_wlook[kMaxT / 32][2];
    __shared__ unsigned long long s_base_total[2];
    __shared__ unsigned long long s_trace[16];
    __shared__ int s_stat[8];

    const int T = blockDim.x;
    const int tid = threadIdx.x;
    const int nwarp = T >> 5;
    const int warp = tid >> 5;
    const unsigned lane = lane_id();
    const int G = gridDim.x;
    const int b = blockIdx.x;
    const FrameParams& P = A.P;
    const int64_t n = P.n;

    // shared-memory carve-up: points | keys | ring (per warp) | group popcounts
    float4* s_pts = reinterpret_cast<float4*>(fsm);
    unsigned* s_key = reinterpret_cast<unsigned*>(fsm + (size_t)A.smem_points * 16);
    uint2* s_ring = reinterpret_cast<uint2*>(fsm + (size_t)A.smem_points * 20);
    unsigned char* s_gcnt = reinterpret_cast<unsigned char*>(s_ring + (size_t)nwarp * kFusedRing);

    // this CTA's chunk of the frame: [c0, c0 + m), the first `res` points resident in shared memory
    const int64_t per = ((n + G - 1) / G + 31) & ~(int64_t)31;   // multiple of 32: whole warps, 512-B aligned
    const int64_t c0 = (int64_t)b * per < n ? (int64_t)b * per : n;
    const int m = (int)((c0 + per <= n ? c0 + per : n) - c0);
    const int res = m < A.smem_points ? m : A.smem_points;
    const float4* gp = P.pts + c0;
    const int stage_pts = ((res + kFusedLoadStages - 1) / kFusedLoadStages + 3) & ~3;

    // Programmatic dependent launch: a frame written by the kernel just before this one on the stream (a crop, a
    // torch.cat, a copy kernel) is only guaranteed visible after griddepcontrol.wait; the load may run ahead of the wait
    // only when the caller vouches that the input was complete before the launch (lidar_frame_set_fused_pdl(2)).
    if (!A.early_load) asm volatile("griddepcontrol.wait;" ::: "memory");
    if (tid == 0) {
        s_trace[0] = global_timer_ns();
#pragma unroll
        for (int s = 0; s < kFusedLoadStages; ++s) fused_mbar_init(&s_bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#pragma unroll
        for (int s = 0; s < kFusedLoadStages; ++s) {
            const int p0 = s * stage_pts;
            int cnt = res - p0;
            cnt = cnt < 0 ? 0 : (cnt > stage_pts ? stage_pts : cnt);
            if (cnt > 0) {
                fused_mbar_expect_tx(&s_bar[s], (unsigned)cnt * 16u);
                fused_bulk_g2s(s_pts + p0, gp + p0, (unsigned)cnt * 16u, &s_bar[s]);
            }
        }
    }
    // let the NEXT frame's kernel (programmatic dependent launch) start as SMs free up: its TMA load and
    // bounding box touch nothing this kernel owns, everything else waits at its griddepcontrol.wait
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    __syncthreads();   // mbarrier init visible to every waiter
    // ---- phase 0c: bounding box of the chunk -----------------------------------------------------
    {
        float mn[4] = {INFINITY, INFINITY, INFINITY, INFINITY};
        float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        auto fold = [&](const float4& v) {
            mn[0] = fminf(mn[0], v.x); mx[0] = fmaxf(mx[0], v.x);
            mn[1] = fminf(mn[1], v.y); mx[1] = fmaxf(mx[1], v.y);
            mn[2] = fminf(mn[2], v.z); mx[2] = fmaxf(mx[2], v.z);
            mn[3] = fminf(mn[3], v.w); mx[3] = fmaxf(mx[3], v.w);
        };
        // spill points first: their global loads overlap the wait for the bulk copies
        for (int j = res + tid; j < m; j += T) fold(__ldcg(gp + j));
#pragma unroll
        for (int s = 0; s < kFusedLoadStages; ++s) {
            const int p0 = s * stage_pts;
            int p1 = p0 + stage_pts;
            p1 = p1 > res ? res : p1;
            if (p0 < p1) {
                fused_mbar_wait(&s_bar[s], 0);
                for (int j = p0 + tid; j < p1; j += T) fold(s_pts[j]);
            }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o));
                mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
            }
        }
        if (lane == 0) {
#pragma unroll
            for (int c = 0; c < 4; ++c) { s_red[warp][c] = (double)mn[c]; s_red[warp][4 + c] = (double)mx[c]; }
        }
        __syncthreads();
    }
    // ---- phase 0b: from here on the frame touches state shared with the previous frame of this pipeline
    //      (workspace, outputs): wait for that kernel to have completed (no-op without a dependent launch).
    //      Then zero the density grid and the scan flags; the occupancy bitmap is already clean unless the
    //      previous frame of this workspace ran on the five-kernel path (dirty > 0) ---------------------
    asm volatile("griddepcontrol.wait;" ::: "memory");
    {
        const int64_t gt0 = (int64_t)b * T + tid, gstride = (int64_t)G * T;
        int64_t dirty = (int64_t)A.ctrl->dirty_groups;
        if (dirty > A.groups_cap) dirty = A.groups_cap;
        uint4* b4 = reinterpret_cast<uint4*>(A.groups);
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        for (int64_t k = gt0; k < dirty * 2; k += gstride) b4[k] = z;
        for (int64_t k = gt0; k < A.grid_cap; k += gstride) A.grid_out[k] = 0;
        for (int64_t k = gt0; k < G; k += gstride) A.cta_desc[k] = 0ull;
    }
    if (tid < 8) {
        const bool is_max = tid >= 4;
        double v = is_max ? -INFINITY : INFINITY;
        for (int w = 0; w < nwarp; ++w) v = is_max ? fmax(v, s_red[w][tid]) : fmin(v, s_red[w][tid]);
        A.partial[(size_t)b * 8 + tid] = v;
    }
    FUSED_TRACE(1);
    fused_grid_barrier(&A.ctrl->grid_bar, 1u * G);
    FUSED_TRACE(2);
    // ---- every CTA folds the G partials and derives the descriptor (identical in every CTA) ------
    {
        const int ch = tid & 7;
        const bool is_max = ch >= 4;
        double v = is_max ? -INFINITY : INFINITY;
        for (int q = tid >> 3; q < G; q += T / 8) {
            const double x = __ldcg(A.partial + (size_t)q * 8 + ch);
            v = is_max ? fmax(v, x) : fmin(v, x);
        }
        for (int o = 8; o < 32; o <<= 1) {
            const double x = __shfl_xor_sync(0xffffffffu, v, o);
            v = is_max ? fmax(v, x) : fmin(v, x);
        }
        if (lane < 8) s_red[warp][lane] = v;   // the barrier above separates this from the earlier use
        __syncthreads();
        if (tid < 8) {
            double r = s_red[0][tid];
            for (int w = 1; w < nwarp; ++w) r = (tid >= 4) ? fmax(r, s_red[w][tid]) : fmin(r, s_red[w][tid]);
            s_bb[tid] = r;
        }
        __syncthreads();
        derive_desc_cta(P, s_bb, &D, s_stat);
    }
    FUSED_TRACE(3);
    const bool ok = D.status == 0;
    const int64_t ng = ok ? (D.key_space + kGroupVoxels - 1) / kGroupVoxels : 0;
    // scan-order variant, sparse key space (a 240 m x 240 m ring scan: 2.8 M groups, 0.3 M of them occupied): scan and
    // clean walk a summary bitmap (one bit per group, set by whoever touches an empty word) instead of streaming every
    // group, so the frame costs what its DATA costs, not what its bounding box costs.  Same decision in every CTA.
    const int64_t l1_used = (ng + 31) / 32;
    const int64_t lpc = (l1_used + G - 1) / G;                 // summary words per CTA (prefix pass)
    const bool sparse = kScan && ok && 2 * ng > 3 * n;
    __shared__ unsigned s_cta_base[kScan ? kFusedMaxCtas : 1];
    // density grid replicas: R copies of nx*ny cells, CTA b adds into copy b % R
    const int ncell = D.nx * D.ny;
    int grid_reps = (ok && ncell > 0) ? kGridRepCells / ncell : 0;
    grid_reps = grid_reps > kGridRepMax ? kGridRepMax : grid_reps;
    int32_t* const grid_acc = grid_reps >= 2 ? A.grid_rep + (size_t)(b % grid_reps) * ncell : A.grid_out;

    // ---- phase 1: mark ----------------------------------------------------------------------------
    if (ok) {
        const MarkConst K = make_mark_const(D);
        if constexpr (kScan) {
            // Run-length aggregation across the warp: adjacent lanes hold adjacent points, and a sensor delivers them
            // in scan order -- 25 consecutive returns of a ring share a 5 cm voxel near the sensor.  Only the first
            // lane of a run of equal keys touches the bitmap (the others are later members by construction), and
            // only the first lane of a run of equal density cells issues the reduction, with the run length: most of
            // the L2 atomics of such a frame disappear, all of them same-address (serialised) ones.
            for (int j0 = tid; j0 - (int)lane < res; j0 += kFusedBatch * T) {      // warp-uniform trip count
                float4 q[kFusedBatch];
                int key[kFusedBatch];
                unsigned bit[kFusedBatch], old[kFusedBatch];
#pragma unroll
                for (int u = 0; u < kFusedBatch; ++u) {
                    const int j = j0 + u * T;
                    q[u] = s_pts[j < res ? j : 0];
                }
#pragma unroll
                for (int u = 0; u < kFusedBatch; ++u) {
                    const bool live = j0 + u * T < res;
                    key[u] = voxel_key_of(q[u], D, K);
                    const int kk = live ? key[u] : -1 - (int)lane;             // dead lanes never join a run
                    const int prev = __shfl_up_sync(0xffffffffu, kk, 1);
                    const bool head = lane == 0 || kk != prev;
                    uint32_t* w = occupancy_word(A.groups, key[u], bit[u]);
                    old[u] = bit[u];               // later members of a run, and dead slots: "bit already set"
                    if (live && head) old[u] = atomicOr(w, bit[u]);
                }
                if (sparse) {
#pragma unroll
                    for (int u = 0; u < kFusedBatch; ++u) {
                        if (old[u] == 0u) {                                    // first to touch this word
                            const unsigned g = (unsigned)key[u] / kGroupVoxels;
                            red_or_b32(A.l1 + (g >> 5), 1u << (g & 31));
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < kFusedBatch; ++u) {
                    const int j = j0 + u * T;
                    const bool live = j < res;
                    if (live) st_stream_s32(A.voxel_key + c0 + j, key[u]);
                    if (K.do_grid) {
                        const int cell = live ? grid_cell_of(q[u], D, K) : -1;
                        const int prev = __shfl_up_sync(0xffffffffu, cell, 1);
                        const bool chead = lane == 0 || cell != prev;
                        const unsigned heads = __ballot_sync(0xffffffffu, chead);
                        if (chead && cell >= 0) {
                            const unsigned above = lane == 31 ? 0u : heads & (0xffffffffu << (lane + 1));
                            red_add_s32(grid_acc + cell, (above ? __ffs(above) - 1 : 32) - (int)lane);
                        }
                    }
                    if (live) s_key[j] = (unsigned)key[u] | ((old[u] & bit[u]) ? kDupFlag : 0u);
                }
            }
        } else {
            // resident points: kFusedBatch points per thread per trip, all their atomicOr in flight together
            for (int j0 = tid; j0 < res; j0 += kFusedBatch * T) {
                float4 q[kFusedBatch];
                int key[kFusedBatch];
                unsigned bit[kFusedBatch], old[kFusedBatch];
#pragma unroll
                for (int u = 0; u < kFusedBatch; ++u) {
                    const int j = j0 + u * T;
                    q[u] = s_pts[j < res ? j : j0];
                }
#pragma unroll
                for (int u = 0; u < kFusedBatch; ++u) {
                    key[u] = voxel_key_of(q[u], D, K);
                    uint32_t* w = occupancy_word(A.groups, key[u], bit[u]);
                    old[u] = 0u;
                    if (j0 + u * T < res) old[u] = atomicOr(w, bit[u]);
                }
#pragma unroll
                for (int u = 0; u < kFusedBatch; ++u) {
                    const int j = j0 + u * T;
                    if (j < res) {
                        st_stream_s32(A.voxel_key + c0 + j, key[u]);
                        if (K.do_grid) {
                            const int cell = grid_cell_of(q[u], D, K);
                            if (cell >= 0) red_add_s32(grid_acc + cell, 1);
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < kFusedBatch; ++u) {
                    const int j = j0 + u * T;
                    if (j < res) s_key[j] = (unsigned)key[u] | ((old[u] & bit[u]) ? kDupFlag : 0u);
                }
            }
        }
        // spill points: the flagged key goes to global memory, phase 3 strips the flag
        for (int j = res + tid; j < m; j += T) {
            const float4 q = __ldcg(gp + j);
            const int key = voxel_key_of(q, D, K);
            unsigned bit;
            uint32_t* w = occupancy_word(A.groups, key, bit);
            const unsigned old = atomicOr(w, bit);
            if (sparse && old == 0u) {
                const unsigned g = (unsigned)key / kGroupVoxels;
                red_or_b32(A.l1 + (g >> 5), 1u << (g & 31));
            }
            if (K.do_grid) {
                const int cell = grid_cell_of(q, D, K);
                if (cell >= 0) red_add_s32(grid_acc + cell, 1);
            }
            A.voxel_key[c0 + j] = (int)((unsigned)key | ((old & bit) ? kDupFlag : 0u));
        }
    }
    FUSED_TRACE(4);
    fused_grid_barrier(&A.ctrl->grid_bar, 2u * G);
    FUSED_TRACE(5);

    // ---- phase 2: scan of the occupancy groups ----------------------------------------------------
    // CTA b owns groups [g0, g1); warp w owns the contiguous segment [w0, w1) and walks it in rows of 32
    // groups (1 KB, one 256-bit load per lane)
    const int64_t gpc = (ng + G - 1) / G;
    const int64_t g0 = (int64_t)b * gpc < ng ? (int64_t)b * gpc : ng;
    const int64_t g1 = g0 + gpc < ng ? g0 + gpc : ng;
    const int64_t wseg = (((gpc + nwarp - 1) / nwarp) + 31) & ~(int64_t)31;
    const int64_t w0 = g0 + (int64_t)warp * wseg < g1 ? g0 + (int64_t)warp * wseg : g1;
    const int64_t w1 = w0 + wseg < g1 ? w0 + wseg : g1;
    unsigned long long n_voxels = 0ull;
    // sparse scan (scan-order variant).  Pass A, summary words dealt round-robin to the CTAs (occupied words cluster
    // in key space -- near the sensor -- so contiguous ranges would leave most CTAs idle): occupied cells per word.  Pass B, contiguous range per CTA: exclusive prefix inside the range, total of the range.
    const int64_t lw0 = (int64_t)b * lpc < l1_used ? (int64_t)b * lpc : l1_used;
    const int64_t lw1 = lw0 + lpc < l1_used ? lw0 + lpc : l1_used;
    if (sparse) {
        // warp gw of the grid takes the words gw, gw + W, ...; each lane fetches one of them, then the warp visits
        // the non-empty ones two at a time with one LANE per occupancy group (64 group loads in flight per warp)
        const int W = G * nwarp, gw = warp * G + b;
        for (int64_t k0 = 0; gw + k0 * W < l1_used; k0 += 32) {
            const int64_t mylw = gw + (k0 + lane) * (int64_t)W;
            const unsigned myword = mylw < l1_used ? __ldcg(A.l1 + mylw) : 0u;
            unsigned nz = __ballot_sync(0xffffffffu, myword != 0u);
            unsigned mycnt = 0;
            while (nz) {
                const int s0 = __ffs(nz) - 1;
                nz &= nz - 1;
                const int s1 = nz ? __ffs(nz) - 1 : s0;
                nz &= nz - 1;                                    // no-op when nz was already 0
                const unsigned w0 = __shfl_sync(0xffffffffu, myword, s0);
                const unsigned w1 = s1 != s0 ? __shfl_sync(0xffffffffu, myword, s1) : 0u;
                const int64_t g0_ = (gw + (k0 + s0) * (int64_t)W) * 32 + lane, g1_ = (gw + (k0 + s1) * (int64_t)W) * 32 + lane;
                unsigned pc0 = 0, pc1 = 0;
                Word8 a, c;
                if ((w0 >> lane) & 1u) a = ld_cg_256(A.groups + (size_t)g0_ * 8);
                if ((w1 >> lane) & 1u) c = ld_cg_256(A.groups + (size_t)g1_ * 8);
                if ((w0 >> lane) & 1u) pc0 = group_popc(a);
                if ((w1 >> lane) & 1u) pc1 = group_popc(c);
                pc0 = warp_sum_u32(pc0);
                pc1 = warp_sum_u32(pc1);
                if ((int)lane == s0) mycnt = pc0;
                if ((int)lane == s1 && s1 != s0) mycnt = pc1;
            }
            if (mylw < l1_used) A.l1cnt[mylw] = mycnt;
        }
        fused_grid_barrier(&A.ctrl->grid_bar_b, (unsigned)G);
        const int64_t lchunk = (lpc + T - 1) / T;
        const int64_t t0w = lw0 + (int64_t)tid * lchunk < lw1 ? lw0 + (int64_t)tid * lchunk : lw1;
        const int64_t t1w = t0w + lchunk < lw1 ? t0w + lchunk : lw1;
        unsigned mine = 0;
        for (int64_t lw = t0w; lw < t1w; ++lw) mine += __ldcg(A.l1cnt + lw);
        unsigned inc = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= (unsigned)o) inc += t;
        }
        if (lane == 31) s_wsum[warp] = inc;
        __syncthreads();
        unsigned total = 0, warp_off = 0;
        for (int w = 0; w < nwarp; ++w) {
            if (w < warp) warp_off += s_wsum[w];
            total += s_wsum[w];
        }
        unsigned run = warp_off + inc - mine;
        for (int64_t lw = t0w; lw < t1w; ++lw) {
            const unsigned c = __ldcg(A.l1cnt + lw);
            A.l1cnt[lw] = run;                  // exclusive prefix inside this CTA's range
            run += c;
        }
        if (tid == 0) A.cta_desc[b] = (unsigned long long)total;
        FUSED_TRACE(6);
    } else if (ok) {
        const bool cache = gpc <= (int64_t)A.smem_groups;
        unsigned mine = 0;
        for (int64_t row = w0; row < w1; row += 32 * kFusedBatch) {
            Word8 gw[kFusedBatch];
#pragma unroll
            for (int u = 0; u < kFusedBatch; ++u) {          // kFusedBatch rows (1 KB each) in flight per warp
                const int64_t g = row + u * 32 + lane;
                if (g < w1) gw[u] = ld_cg_256(A.groups + (size_t)g * 8);
            }
#pragma unroll
            for (int u = 0; u < kFusedBatch; ++u) {
                const int64_t g = row + u * 32 + lane;
                if (g < w1) {
                    const unsigned pc = group_popc(gw[u]);
                    if (cache) s_gcnt[g - g0] = (unsigned char)pc;
                    mine += pc;
                }
            }
        }
        mine = warp_sum_u32(mine);
        if (lane == 0) s_wsum[warp] = mine;
        __syncthreads();
        unsigned total = 0;
        for (int w = 0; w < nwarp; ++w) total += s_wsum[w];
        if (tid == 0) A.cta_desc[b] = (unsigned long long)total;
        FUSED_TRACE(6);
    }
    // totals of ALL CTAs: earlier ones give this CTA's base, all of them give the voxel count.  (One
    // poller per CTA on the barrier counter; polling the G totals directly puts G*G threads on a few lines.)
    fused_grid_barrier(&A.ctrl->grid_bar, 3u * G);
    if (ok) {
        const bool cache = gpc <= (int64_t)A.smem_groups;
        unsigned long long before = 0ull, all = 0ull;
        for (int c = tid; c < G; c += T) {
            const unsigned long long d = __ldcg(A.cta_desc + c);
            all += d;
            if (c < b) before += d;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            before += __shfl_xor_sync(0xffffffffu, before, o);
            all += __shfl_xor_sync(0xffffffffu, all, o);
        }
        if (lane == 0) { s_wlook[warp][0] = before; s_wlook[warp][1] = all; }
        __syncthreads();
        if (tid == 0) {
            unsigned long long bs = 0ull, al = 0ull;
            for (int w = 0; w < nwarp; ++w) { bs += s_wlook[w][0]; al += s_wlook[w][1]; }
            s_base_total[0] = bs;
            s_base_total[1] = al;
            D.n_voxels = (int64_t)al;
        }
        __syncthreads();
        FUSED_TRACE(7);
        n_voxels = s_base_total[1];
        if (sparse) {
            // pass C, strided again: base of the owning range + prefix inside it, then the prefix of every occupied
            // group of the word goes into word 0 of the group
            {
                unsigned long long acc = 0ull;          // exclusive scan of the G range totals (G <= 1024: one warp)
                if (warp == 0) {
                    for (int c0_ = 0; c0_ < G; c0_ += 32) {
                        const int c = c0_ + (int)lane;
                        unsigned v = c < G ? (unsigned)__ldcg(A.cta_desc + c) : 0u;
                        unsigned inc = v;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
                            if (lane >= (unsigned)o) inc += t;
                        }
                        if (c < G) s_cta_base[c] = (unsigned)acc + inc - v;
                        acc += __shfl_sync(0xffffffffu, inc, 31);
                    }
                }
                __syncthreads();
            }
            const int W = G * nwarp, gw = warp * G + b;
            for (int64_t k0 = 0; gw + k0 * W < l1_used; k0 += 32) {
                const int64_t mylw = gw + (k0 + lane) * (int64_t)W;
                const unsigned myword = mylw < l1_used ? __ldcg(A.l1 + mylw) : 0u;
                const unsigned mybase = myword ? s_cta_base[(int)(mylw / lpc)] + __ldcg(A.l1cnt + mylw) : 0u;
                unsigned nz = __ballot_sync(0xffffffffu, myword != 0u);
                while (nz) {
                    const int s0 = __ffs(nz) - 1;
                    nz &= nz - 1;
                    const int s1 = nz ? __ffs(nz) - 1 : s0;
                    nz &= nz - 1;
                    const unsigned w0 = __shfl_sync(0xffffffffu, myword, s0);
                    const unsigned w1 = s1 != s0 ? __shfl_sync(0xffffffffu, myword, s1) : 0u;
                    const unsigned b0 = __shfl_sync(0xffffffffu, mybase, s0), b1 = __shfl_sync(0xffffffffu, mybase, s1);
                    const int64_t g0_ = (gw + (k0 + s0) * (int64_t)W) * 32 + lane, g1_ = (gw + (k0 + s1) * (int64_t)W) * 32 + lane;
                    const bool on0 = (w0 >> lane) & 1u, on1 = (w1 >> lane) & 1u;
                    Word8 a, c;
                    if (on0) a = ld_cg_256(A.groups + (size_t)g0_ * 8);
                    if (on1) c = ld_cg_256(A.groups + (size_t)g1_ * 8);
                    const unsigned pc0 = on0 ? group_popc(a) : 0u, pc1 = on1 ? group_popc(c) : 0u;
                    unsigned i0 = pc0, i1 = pc1;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const unsigned t0 = __shfl_up_sync(0xffffffffu, i0, o), t1 = __shfl_up_sync(0xffffffffu, i1, o);
                        if (lane >= (unsigned)o) { i0 += t0; i1 += t1; }
                    }
                    if (on0) A.groups[(size_t)g0_ * 8] = b0 + i0 - pc0;
                    if (on1) A.groups[(size_t)g1_ * 8] = b1 + i1 - pc1;
                }
            }
        }
        unsigned warp_off = 0;
        for (int w = 0; w < warp; ++w) warp_off += s_wsum[w];
        unsigned run = (unsigned)s_base_total[0] + warp_off;
        for (int64_t row = w0; row < (sparse ? w0 : w1); row += 32) {
            const int64_t g = row + lane;
            unsigned pc = 0;
            if (g < w1) pc = cache ? (unsigned)s_gcnt[g - g0] : group_popc(ld_cg_256(A.groups + (size_t)g * 8));
            unsigned inc = pc;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= (unsigned)o) inc += t;
            }
            if (g < w1) A.groups[(size_t)g * 8] = run + inc - pc;
            run += __shfl_sync(0xffffffffu, inc, 31);
        }
    }
    FUSED_TRACE(8);
    fused_grid_barrier(&A.ctrl->grid_bar, 4u * G);
    FUSED_TRACE(9);

    // ---- phase 3: rank ----------------------------------------------------------------------------
    if (ok) {
        uint2* ring = s_ring + (size_t)warp * kFusedRing;
        unsigned head = 0, count = 0;   // warp-uniform
        // Drains up to 32 parked later-members (all lanes enter; `active` is false in the tail of the last drain).
        // Scan-order variant: consecutive entries of the ring are consecutive points, which in scan order belong to the
        // same voxel; a segmented inclusive scan over runs of equal rank (5 shuffle steps) leaves the exact integer
        // sum of every run in its last lane, which issues the five reductions for the whole run.
        auto drain = [&](unsigned slot, bool active) {
            if constexpr (kScan) {
                const uint2 e = active ? ring[slot] : make_uint2(0u, 0xffffffffu - lane);   // inactive: a run of its own
                Fixed4 f = {0, 0, 0, 0};
                if (active) {
                    const int j = (int)e.x;
                    float4 q;
                    int key;
                    if (j < res) { q = s_pts[j]; key = (int)(s_key[j] & ~kDupFlag); }
                    else { q = __ldcg(gp + j); key = __ldcg(A.voxel_key + c0 + j); }   // flag already stripped
                    f = fixed_offsets(q, key, D);
                }
                const unsigned r = e.y;
                const unsigned prev = __shfl_up_sync(0xffffffffu, r, 1);
                const unsigned heads = __ballot_sync(0xffffffffu, lane == 0 || r != prev);
                const int start = 31 - __clz((int)(heads & (0xffffffffu >> (31 - lane))));   // first lane of this run
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const long long x = __shfl_up_sync(0xffffffffu, f.x, o), y = __shfl_up_sync(0xffffffffu, f.y, o);
                    const long long z = __shfl_up_sync(0xffffffffu, f.z, o), w = __shfl_up_sync(0xffffffffu, f.w, o);
                    if ((int)lane - o >= start) { f.x += x; f.y += y; f.z += z; f.w += w; }
                }
                const bool tail = lane == 31 || ((heads >> (lane + 1)) & 1u);
                if (active && tail) {
                    unsigned long long* Acc = reinterpret_cast<unsigned long long*>(A.acc + (size_t)r * 4);
                    red_add_u64(Acc + 0, (unsigned long long)f.x);
                    red_add_u64(Acc + 1, (unsigned long long)f.y);
                    red_add_u64(Acc + 2, (unsigned long long)f.z);
                    red_add_u64(Acc + 3, (unsigned long long)f.w);
                    red_add_s32(A.cnt + r, (int)lane - start + 1);
                }
            } else {
                if (!active) return;
                const uint2 e = ring[slot];
                const int j = (int)e.x;
                float4 q;
                int key;
                if (j < res) { q = s_pts[j]; key = (int)(s_key[j] & ~kDupFlag); }
                else { q = __ldcg(gp + j); key = __ldcg(A.voxel_key + c0 + j); }   // flag already stripped
                accumulate_dup(q, key, e.y, D, A.acc, A.cnt);
            }
        };
        const int m_round = (m + 31) & ~31;
        for (int j0 = tid; j0 < m_round; j0 += kFusedBatch * T) {
            float4 q[kFusedBatch];
            unsigned fk[kFusedBatch], r[kFusedBatch];
            bool live[kFusedBatch];
#pragma unroll
            for (int u = 0; u < kFusedBatch; ++u) {
                const int j = j0 + u * T;
                live[u] = j < m;
                fk[u] = 0u;
                r[u] = 0u;
                if (live[u]) {
                    if (j < res) { q[u] = s_pts[j]; fk[u] = s_key[j]; }
                    else { q[u] = __ldcg(gp + j); fk[u] = (unsigned)__ldcg(A.voxel_key + c0 + j); }
                }
            }
#pragma unroll
            for (int u = 0; u < kFusedBatch; ++u)
                if (live[u]) r[u] = rank_of((int)(fk[u] & ~kDupFlag), A.groups);
#pragma unroll
            for (int u = 0; u < kFusedBatch; ++u) {
                const int j = j0 + u * T;
                const bool dup = live[u] && (fk[u] & kDupFlag) != 0u;
                if (live[u]) {
                    const int key = (int)(fk[u] & ~kDupFlag);
                    st_stream_s32(A.inverse + c0 + j, (int)r[u]);
                    if (!dup) store_first_member(A.voxels, r[u], q[u], key);
                    else if (j >= res) A.voxel_key[c0 + j] = key;      // strip the flag of a spilled key
                }
                const unsigned mm = __ballot_sync(0xffffffffu, dup);
                if (mm) {
                    if (dup) ring[(head + count + __popc(mm & lanemask_lt())) % kFusedRing] = make_uint2((unsigned)j, r[u]);
                    count += __popc(mm);
                    __syncwarp();
                    if (count >= 32) {
                        drain((head + lane) % kFusedRing, true);
                        head = (head + 32) % kFusedRing;
                        count -= 32;
                        __syncwarp();
                    }
                }
            }
        }
        if (count) drain((head + lane) % kFusedRing, lane < count);
    }
    FUSED_TRACE(10);
    fused_grid_barrier(&A.ctrl->grid_bar, 5u * G);
    FUSED_TRACE(11);

    // ---- phase 4: clean the bitmap for the next frame, finalize the multi-member voxels -----------
    if (grid_reps >= 2) {
        // merge the density replicas (complete since barrier 2) into the output grid and leave them zeroed
        // for the next frame; placed here because this phase is latency-bound and has the slack
        for (int k = b * T + tid; k < ncell; k += G * T) {
            int v[kGridRepMax];
#pragma unroll
            for (int r = 0; r < kGridRepMax; ++r)      // all replica loads in flight before the first store
                v[r] = r < grid_reps ? __ldcg(A.grid_rep + (size_t)r * ncell + k) : 0;
            int sum = 0;
#pragma unroll
            for (int r = 0; r < kGridRepMax; ++r) {
                sum += v[r];
                if (r < grid_reps) A.grid_rep[(size_t)r * ncell + k] = 0;
            }
            A.grid_out[k] = sum;
        }
    }
    if (ok) {
        Word8 z;
#pragma unroll
        for (int k = 0; k < 8; ++k) z.w[k] = 0u;
        if (sparse) {
            for (int64_t lw = (int64_t)tid * G + b; lw < l1_used; lw += (int64_t)G * T) {   // word w -> CTA w % G
                unsigned word = __ldcg(A.l1 + lw);
                if (word) A.l1[lw] = 0u;
                while (word) {
                    const int64_t g = lw * 32 + (__ffs(word) - 1);
                    word &= word - 1;
                    st_256(A.groups + (size_t)g * 8, z);
                }
            }
        } else {
            for (int64_t row = w0; row < w1; row += 32) {
                const int64_t g = row + lane;
                if (g < w1) st_256(A.groups + (size_t)g * 8, z);
            }
        }
        unsigned* ring = reinterpret_cast<unsigned*>(s_ring + (size_t)warp * kFusedRing);
        unsigned head = 0, count = 0;
        const double isx = 1.0 / D.fix_scale_xyz, isw = 1.0 / D.fix_scale_w;
        const int64_t V = (int64_t)n_voxels;
        const int64_t vper = ((V + G - 1) / G + 31) & ~(int64_t)31;
        const int64_t v0 = (int64_t)b * vper;
        const int64_t v1 = v0 + vper;   // whole warps; bounds-checked against V below
        for (int64_t r0 = v0 + tid; r0 < v1; r0 += (int64_t)kFusedBatch * T) {
            int c[kFusedBatch];
#pragma unroll
            for (int u = 0; u < kFusedBatch; ++u) {
                const int64_t r = r0 + (int64_t)u * T;
                c[u] = (r < v1 && r < V) ? __ldcg(A.cnt + r) : 0;
            }
#pragma unroll
            for (int u = 0; u < kFusedBatch; ++u) {
                const int64_t r = r0 + (int64_t)u * T;
                const bool hit = c[u] != 0;
                const unsigned mm = __ballot_sync(0xffffffffu, hit);
                if (mm) {
                    if (hit) ring[(head + count + __popc(mm & lanemask_lt())) % kFusedRing] = (unsigned)r;
                    count += __popc(mm);
                    __syncwarp();
                    if (count >= 32) {
                        finalize_voxel(ring[(head + lane) % kFusedRing], D, isx, isw, A.acc, A.cnt, A.voxels);
                        head = (head + 32) % kFusedRing;
                        count -= 32;
                        __syncwarp();
                    }
                }
            }
        }
        if (lane < count) finalize_voxel(ring[(head + lane) % kFusedRing], D, isx, isw, A.acc, A.cnt, A.voxels);
    }
    // ---- epilogue: descriptor out, counters reset by the last CTA to leave -------------------------
    __syncthreads();
    if (tid == 0) {
        s_trace[12] = global_timer_ns();
        for (int k = 0; k < 13; ++k) A.trace_all[(size_t)b * 16 + k] = s_trace[k];
        if (b == 0) {
#pragma unroll
            for (int k = 0; k < 12; ++k) D.trace_ns[k] = (uint32_t)(s_trace[k + 1] - s_trace[k]);
            D.trace_ns[14] = sparse ? 1u : 0u;
            D.trace_ns[15] = (uint32_t)G;
            *A.D = D;
        }
        __threadfence();
        if (atomicAdd(&A.ctrl->exit_ticket, 1u) == (unsigned)G - 1u) {
            A.ctrl->grid_bar = 0u;
            A.ctrl->grid_bar_b = 0u;
            A.ctrl->exit_ticket = 0u;
            A.ctrl->dirty_groups = 0ull;     // phase 4 left the bitmap clean
        }
    }
}

// ================================================================================================
// k_frame_part — the frame as ONE persistent kernel around an MSD radix PARTITION by voxel-key range
// (the north star's "radix-sort-by-voxel-key", done as a single most-significant-digit pass).
//
// k_frame_fused pays four scattered 32-byte L2 operations per point on shuffled input (occupancy atomic,
// density reduction, group load, record store) plus a 22 MB bitmap that is streamed twice.  Here the only
// per-point traffic that is scattered by nature remains: each point's 32-byte record has to land at its
// voxel's rank.  Everything else becomes shared-memory work or run-coalesced traffic:
//
//   phase 0  as k_frame_fused: TMA load of the CTA's chunk into shared memory, bounding box, descriptor
//   phase 1  per point: voxel key, density cell, partition = key >> 18 (2^18 cells = 32 KB of occupancy
//            bits); a shared-memory histogram hands every point its slot inside the CTA's run for that
//            partition; ONE global atomicAdd per (CTA, partition) claims the run's place in the bucket
//   phase 2  exclusive scan of the partition totals (every CTA, redundantly: <= 2048 values), CTA-local
//            counting sort by partition (16-bit permutation in shared memory), then the 8-byte entries
//            {key, cell} go to the bucket in SORTED order: adjacent threads write adjacent addresses
//   phase 3  owner pass: CTA b owns a contiguous range of partitions, one at a time with all threads: occupancy
//            bits + popcount prefix live in shared memory (atomicOr returns "first member or later one"), the
//            density cells of the slab are counted in shared memory and flushed once; per entry the owner
//            returns {rank inside the partition, later-member flag}; per partition its voxel count
//   phase 4  exclusive scan of the partition voxel counts; every CTA reads the answers for ITS points back
//            (sorted order: coalesced runs), inverse goes out coalesced, first members store their record,
//            later members accumulate exact fixed-point sums (same device functions as k_frame_fused)
//   phase 5  finalize the multi-member voxels (sweep of the member counters)
//
// No occupancy bitmap in global memory (nothing to stream, nothing to clean), no density replicas.
// Outputs are identical to the other back ends (tests/test_gpu_frame_modes.py).  Eligible when the whole chunk of
// every CTA is resident in shared memory and caps.max_key_space <= 2^29 (<= 2048 partitions).
// ================================================================================================
constexpr int kPartShift = 18;
constexpr int kPartCells = 1 << kPartShift;
constexpr int kPartWords64 = kPartCells / 64;        // 4096 prefix slots: one per 64 occupancy bits (two 32-bit words)
constexpr int kPartMax = 2048;                       // partitions per frame
constexpr int kPartDensCap = 1024;                   // density cells of a CTA's slab counted in shared memory
constexpr unsigned kPartDup = 0x80000000u;
constexpr int kPartRegs = 8;                        // bucket entries a thread of the owner keeps in registers per partition

struct PartArgs {
    FusedArgs F;
    unsigned* ptotal;     // [kPartMax] entries claimed per partition (all-zero between frames)
    unsigned* pvox;       // [kPartMax] voxels per partition
    uint2* bucket;        // [max_points] {key, density cell} grouped by partition
    unsigned* prank;      // [max_points] owner's answer per bucket slot: rank inside the partition | later-member flag
    int pcap;             // capacity of the per-partition shared-memory arrays (>= partitions of any frame)
    int union_bytes;      // shared-memory region shared by {key, cell, slot} (phases 1-2), the owner structures (3), inverse (4)
};

__device__ __forceinline__ unsigned cta_exclusive_scan_u32(unsigned* a, int n, unsigned* s_wsum /* [nwarp] */, int T) {
    // in-place exclusive scan of a[0..n) by the whole CTA; returns the total.  Each thread owns a contiguous slice.
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = T >> 5;
    const int per = (n + T - 1) / T;
    const int i0 = tid * per < n ? tid * per : n, i1 = i0 + per < n ? i0 + per : n;
    unsigned mine = 0;
    for (int i = i0; i < i1; ++i) mine += a[i];
    unsigned inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_wsum[warp] = inc;
    __syncthreads();
    unsigned woff = 0, total = 0;
    for (int w = 0; w < nwarp; ++w) {
        const unsigned v = s_wsum[w];
        if (w < warp) woff += v;
        total += v;
    }
    unsigned run = woff + inc - mine;
    for (int i = i0; i < i1; ++i) {
        const unsigned v = a[i];
        a[i] = run;
        run += v;
    }
    __syncthreads();
    return total;
}

__global__ void __launch_bounds__(kFusedMaxThreads, 1)
k_frame_part(const PartArgs PA) {
    extern __shared__ __align__(128) unsigned char fsm[];
    __shared__ lidar_frame_desc D;
    __shared__ __align__(8) unsigned long long s_bar[kFusedLoadStages];
    __shared__ double s_red[kFusedMaxThreads / 32][8];
    __shared__ double s_bb[8];
    __shared__ unsigned s_wsum[kFusedMaxThreads / 32];
    __shared__ unsigned long long s_trace[16];
    __shared__ int s_stat[8];

    const FusedArgs& A = PA.F;
    const int T = blockDim.x, tid = threadIdx.x, nwarp = T >> 5, warp = tid >> 5;
    const unsigned lane = lane_id();
    const int G = gridDim.x, b = blockIdx.x;
    const FrameParams& P = A.P;
    const int64_t n = P.n;

    // shared-memory carve-up
    const int spts = A.smem_points;
    float4* s_pts = reinterpret_cast<float4*>(fsm);
    unsigned short* s_perm = reinterpret_cast<unsigned short*>(fsm + (size_t)spts * 16);
    unsigned* s_key = reinterpret_cast<unsigned*>(fsm + (size_t)spts * 18);           // spts is a multiple of 32
    unsigned* s_hist = s_key + spts;
    unsigned* s_delta = s_hist + PA.pcap;
    uint2* s_ring = reinterpret_cast<uint2*>(s_delta + PA.pcap);
    const size_t ring_bytes = (size_t)nwarp * kFusedRing * 8 > 8192 ? (size_t)nwarp * kFusedRing * 8 : 8192;   // doubles as s_pbase[2048]
    unsigned char* s_union = reinterpret_cast<unsigned char*>(s_ring) + ring_bytes;
    // phases 1-2 view of the union
    unsigned short* s_cell = reinterpret_cast<unsigned short*>(s_union);
    unsigned short* s_slot = reinterpret_cast<unsigned short*>(s_union + (size_t)spts * 2);
    // phase 3 view: 2^18 occupancy bits as 32-bit words (native shared-memory atomicOr), a popcount prefix per 64 bits,
    // the density cells of the slab
    unsigned* s_bits = reinterpret_cast<unsigned*>(s_union);                           // 2 * kPartWords64 x 4 B = 32 KB
    unsigned* s_pref = s_bits + 2 * kPartWords64;                                      // 16 KB
    unsigned* s_dens = s_pref + kPartWords64;                                          // kPartDensCap x 4 B
    // phase 4 view
    unsigned* s_inv = reinterpret_cast<unsigned*>(s_union);

    const int64_t per = ((n + G - 1) / G + 31) & ~(int64_t)31;
    const int64_t c0 = (int64_t)b * per < n ? (int64_t)b * per : n;
    const int m = (int)((c0 + per <= n ? c0 + per : n) - c0);       // host guarantees m <= spts
    const float4* gp = P.pts + c0;
    const int stage_pts = ((m + kFusedLoadStages - 1) / kFusedLoadStages + 3) & ~3;

    if (!A.early_load) asm volatile("griddepcontrol.wait;" ::: "memory");
    if (tid == 0) {
        s_trace[0] = global_timer_ns();
#pragma unroll
        for (int s = 0; s < kFusedLoadStages; ++s) fused_mbar_init(&s_bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#pragma unroll
        for (int s = 0; s < kFusedLoadStages; ++s) {
            const int p0 = s * stage_pts;
            int cnt = m - p0;
            cnt = cnt < 0 ? 0 : (cnt > stage_pts ? stage_pts : cnt);
            if (cnt > 0) {
                fused_mbar_expect_tx(&s_bar[s], (unsigned)cnt * 16u);
                fused_bulk_g2s(s_pts + p0, gp + p0, (unsigned)cnt * 16u, &s_bar[s]);
            }
        }
    }
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    for (int p = tid; p < PA.pcap; p += T) s_hist[p] = 0u;
    __syncthreads();
    // ---- phase 0: bounding box of the chunk ----------------------------------------------------------
    {
        float mn[4] = {INFINITY, INFINITY, INFINITY, INFINITY};
        float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int s = 0; s < kFusedLoadStages; ++s) {
            const int p0 = s * stage_pts;
            int p1 = p0 + stage_pts;
            p1 = p1 > m ? m : p1;
            if (p0 < p1) {
                fused_mbar_wait(&s_bar[s], 0);
                for (int j = p0 + tid; j < p1; j += T) {
                    const float4 v = s_pts[j];
                    mn[0] = fminf(mn[0], v.x); mx[0] = fmaxf(mx[0], v.x);
                    mn[1] = fminf(mn[1], v.y); mx[1] = fmaxf(mx[1], v.y);
                    mn[2] = fminf(mn[2], v.z); mx[2] = fmaxf(mx[2], v.z);
                    mn[3] = fminf(mn[3], v.w); mx[3] = fmaxf(mx[3], v.w);
                }
            }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o));
                mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
            }
        }
        if (lane == 0) {
#pragma unroll
            for (int c = 0; c < 4; ++c) { s_red[warp][c] = (double)mn[c]; s_red[warp][4 + c] = (double)mx[c]; }
        }
        __syncthreads();
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    {
        // the occupancy bitmap of the other back ends may be dirty if the previous frame of this workspace ran on the
        // five-kernel path: clean it here so that the workspace invariants hold whichever back end comes next
        const int64_t gt0 = (int64_t)b * T + tid, gstride = (int64_t)G * T;
        int64_t dirty = (int64_t)A.ctrl->dirty_groups;
        if (dirty > A.groups_cap) dirty = A.groups_cap;
        uint4* b4 = reinterpret_cast<uint4*>(A.groups);
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        for (int64_t k = gt0; k < dirty * 2; k += gstride) b4[k] = z;
        for (int64_t k = gt0; k < A.grid_cap; k += gstride) A.grid_out[k] = 0;
    }
    if (tid < 8) {
        const bool is_max = tid >= 4;
        double v = is_max ? -INFINITY : INFINITY;
        for (int w = 0; w < nwarp; ++w) v = is_max ? fmax(v, s_red[w][tid]) : fmin(v, s_red[w][tid]);
        A.partial[(size_t)b * 8 + tid] = v;
    }
    FUSED_TRACE(1);
    fused_grid_barrier(&A.ctrl->grid_bar, 1u * G);
    FUSED_TRACE(2);
    {
        const int ch = tid & 7;
        const bool is_max = ch >= 4;
        double v = is_max ? -INFINITY : INFINITY;
        for (int q = tid >> 3; q < G; q += T / 8) {
            const double x = __ldcg(A.partial + (size_t)q * 8 + ch);
            v = is_max ? fmax(v, x) : fmin(v, x);
        }
        for (int o = 8; o < 32; o <<= 1) {
            const double x = __shfl_xor_sync(0xffffffffu, v, o);
            v = is_max ? fmax(v, x) : fmin(v, x);
        }
        if (lane < 8) s_red[warp][lane] = v;
        __syncthreads();
        if (tid < 8) {
            double r = s_red[0][tid];
            for (int w = 1; w < nwarp; ++w) r = (tid >= 4) ? fmax(r, s_red[w][tid]) : fmin(r, s_red[w][tid]);
            s_bb[tid] = r;
        }
        __syncthreads();
        derive_desc_cta(P, s_bb, &D, s_stat);
    }
    FUSED_TRACE(3);
    const bool ok = D.status == 0;
    const int nparts = ok ? (int)((D.key_space + kPartCells - 1) >> kPartShift) : 0;     // <= pcap (host checked the capacity)
    const int ncell = D.nx * D.ny;
    const bool cell16 = ncell <= 0xffff;           // density cells travel as 16 bits in shared memory; else recomputed
    const MarkConst K = make_mark_const(D);

    // ---- phase 1: keys, cells, slots inside the CTA's per-partition runs ---------------------------------
    if (ok) {
        for (int j0 = tid; j0 < m; j0 += kFusedBatch * T) {
            float4 q[kFusedBatch];
            int key[kFusedBatch], cell[kFusedBatch];
#pragma unroll
            for (int u = 0; u < kFusedBatch; ++u) {
                const int j = j0 + u * T;
                q[u] = s_pts[j < m ? j : j0];
            }
#pragma unroll
            for (int u = 0; u < kFusedBatch; ++u) {
                key[u] = voxel_key_of(q[u], D, K);
                cell[u] = (K.do_grid && cell16) ? grid_cell_of(q[u], D, K) + 1 : 0;          // 0 = outside
            }
#pragma unroll
            for (int u = 0; u < kFusedBatch; ++u) {
                const int j = j0 + u * T;
                if (j < m) {
                    const unsigned slot = atomicAdd(&s_hist[key[u] >> kPartShift], 1u);
                    s_key[j] = (unsigned)key[u];
                    s_slot[j] = (unsigned short)slot;
                    s_cell[j] = (unsigned short)cell[u];
                    st_stream_s32(A.voxel_key + c0 + j, key[u]);
                }
            }
        }
    }
    __syncthreads();
    // claim the place of every non-empty run in its partition's bucket: delta = offset inside the partition
    if (ok) {
        for (int p = tid; p < nparts; p += T) {
            const unsigned c = s_hist[p];
            s_delta[p] = c ? atomicAdd(PA.ptotal + p, c) : 0u;
        }
    }
    FUSED_TRACE(4);
    fused_grid_barrier(&A.ctrl->grid_bar, 2u * G);
    FUSED_TRACE(5);

    // ---- phase 2: partition bases, local counting sort, sorted scatter ------------------------------------
    // owner of partition p: CTA floor(p * G / nparts), i.e. CTA b owns [ceil(b*nparts/G), ceil((b+1)*nparts/G)) -- sizes
    // differ by at most one, every CTA gets work when nparts >= G
    const int p0 = (int)(((long long)b * nparts + G - 1) / G);
    const int p1 = (int)(((long long)(b + 1) * nparts + G - 1) / G);
    unsigned* s_pbase = reinterpret_cast<unsigned*>(s_ring);     // [nparts <= 2048] bucket start of every partition: the
                                                                  // ring area (8 B x 64 x warps >= 8 KB) is idle until phase 4
    unsigned bucket_total = 0;
    if (ok) {
        for (int p = tid; p < nparts; p += T) s_pbase[p] = __ldcg(PA.ptotal + p);
        __syncthreads();
        bucket_total = cta_exclusive_scan_u32(s_pbase, nparts, s_wsum, T);        // == n
        (void)cta_exclusive_scan_u32(s_hist, nparts, s_wsum, T);                   // local sorted position of every run
        for (int p = tid; p < nparts; p += T) s_delta[p] += s_pbase[p];            // bucket position of this CTA's run
        // permutation: sorted position -> point
        for (int j = tid; j < m; j += T) s_perm[s_hist[s_key[j] >> kPartShift] + s_slot[j]] = (unsigned short)j;
        __syncthreads();
        // sorted scatter: thread t writes the entry of sorted position t; a run of one partition is contiguous in the bucket
        // (every CTA starts at a different sorted position: without the rotation all 148 CTAs would write partition 0's
        // window of the bucket at the same time, then partition 1's, ... one hot L2 region after the other)
        const int rot = m ? (int)(((long long)b * m / G) & ~31ll) : 0;
        for (int i = tid; i < m; i += T) {
            const int t = i + rot < m ? i + rot : i + rot - m;
            const int j = s_perm[t];
            const unsigned key = s_key[j];
            const unsigned p = key >> kPartShift;
            const unsigned dest = s_delta[p] + ((unsigned)t - s_hist[p]);
            PA.bucket[dest] = make_uint2(key, (unsigned)s_cell[j]);
        }
        __syncthreads();
        // from here on: bucket slot of sorted position t = s_delta[p] + t
        for (int p = tid; p < nparts; p += T) s_delta[p] -= s_hist[p];
    }
    FUSED_TRACE(6);
    fused_grid_barrier(&A.ctrl->grid_bar, 3u * G);
    FUSED_TRACE(7);

    // ---- phase 3: owner pass over partitions [p0, p1) ------------------------------------------------------
    if (ok) {
        // density cells this CTA's slab can touch: [cell_lo, cell_lo + kPartDensCap); anything else goes straight to
        // global memory (correct whatever the estimate: it only decides where the count is accumulated)
        int cell_lo = 0;
        if (K.do_grid && p0 < p1) {
            const long long cells_per_ix = (long long)D.dims[1] * D.dims[2];
            const int ix0 = (int)(((long long)p0 << kPartShift) / cells_per_ix);
            const double x0 = __dadd_rn(D.origin[0], __dmul_rn((double)ix0, D.voxel));
            int bx = arange_bin(x0, D.ex0, D.ex1, D.exd, K.rdx, D.nx);
            bx = bx < 1 ? 0 : bx - 1;
            cell_lo = bx * D.ny;
        }
        if (K.do_grid) {
            for (int k = tid; k < kPartDensCap; k += T) s_dens[k] = 0u;
        }
        for (int p = p0; p < p1; ++p) {
            const unsigned e0 = s_pbase[p], e1 = p + 1 < nparts ? s_pbase[p + 1] : bucket_total;
            const unsigned kbase = (unsigned)p << kPartShift;
            // the thread's entries stay in registers from the first pass to the second (all loads in flight at once);
            // a partition with more than kPartRegs * T entries takes the streaming loops below for the remainder
            uint2 en[kPartRegs];
            unsigned ans[kPartRegs];
#pragma unroll
            for (int r = 0; r < kPartRegs; ++r) {
                const unsigned e = e0 + tid + r * T;
                en[r] = make_uint2(0u, 0u);
                if (e < e1) en[r] = __ldcg(PA.bucket + e);
            }
            for (int w = tid; w < kPartWords64; w += T) reinterpret_cast<uint2*>(s_bits)[w] = make_uint2(0u, 0u);
            __syncthreads();
            auto mark_one = [&](const uint2& e_) -> unsigned {
                const unsigned local = e_.x - kbase;
                const unsigned bit = 1u << (local & 31);
                const unsigned old = atomicOr(&s_bits[local >> 5], bit);
                if (K.do_grid && cell16) {
                    const int cell = (int)e_.y - 1;
                    if (cell >= 0) {
                        const int rel = cell - cell_lo;
                        if (rel >= 0 && rel < kPartDensCap) atomicAdd(&s_dens[rel], 1u);
                        else red_add_s32(A.grid_out + cell, 1);
                    }
                }
                return (old & bit) ? kPartDup : 0u;
            };
            // pass A: occupancy bits; the old word says whether this entry is the first member of its voxel
#pragma unroll
            for (int r = 0; r < kPartRegs; ++r)
                if (e0 + tid + r * T < e1) ans[r] = mark_one(en[r]);
            for (unsigned e = e0 + tid + kPartRegs * T; e < e1; e += T) PA.prank[e] = mark_one(__ldcg(PA.bucket + e));
            __syncthreads();
            // popcount prefix per 64 bits.  Warp w owns the contiguous words [w*wpw, (w+1)*wpw); in round k its lanes read
            // word base + 32k + lane (consecutive addresses: no bank conflicts) and a warp scan orders them
            {
                const unsigned long long* bits64 = reinterpret_cast<const unsigned long long*>(s_bits);
                const int wpw = kPartWords64 / nwarp;                    // nwarp is a power of two <= 16: 256.. words
                const int base = warp * wpw;
                unsigned wsum = 0;
                for (int k = 0; k < wpw; k += 32) wsum += __popcll(bits64[base + k + lane]);
                wsum = warp_sum_u32(wsum);
                if (lane == 0) s_wsum[warp] = wsum;
                __syncthreads();
                unsigned run = 0, total = 0;
                for (int w = 0; w < nwarp; ++w) { const unsigned v = s_wsum[w]; if (w < warp) run += v; total += v; }
                for (int k = 0; k < wpw; k += 32) {
                    const unsigned pc = __popcll(bits64[base + k + lane]);
                    unsigned inc = pc;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
                        if ((int)lane >= o) inc += t;
                    }
                    s_pref[base + k + lane] = run + inc - pc;
                    run += __shfl_sync(0xffffffffu, inc, 31);
                }
                if (tid == 0) PA.pvox[p] = total;
                __syncthreads();
            }
            // pass B: rank inside the partition
            auto rank_one = [&](unsigned key_) -> unsigned {
                const unsigned local = key_ - kbase;
                const unsigned long long w64 = reinterpret_cast<const unsigned long long*>(s_bits)[local >> 6];
                return s_pref[local >> 6] + __popcll(w64 & ((1ull << (local & 63)) - 1ull));
            };
#pragma unroll
            for (int r = 0; r < kPartRegs; ++r) {
                const unsigned e = e0 + tid + r * T;
                if (e < e1) PA.prank[e] = ans[r] | rank_one(en[r].x);
            }
            for (unsigned e = e0 + tid + kPartRegs * T; e < e1; e += T) PA.prank[e] |= rank_one(__ldcg(&PA.bucket[e].x));
            __syncthreads();
        }
        if (K.do_grid) {
            for (int k = tid; k < kPartDensCap; k += T) {
                const unsigned c = s_dens[k];
                if (c) red_add_s32(A.grid_out + cell_lo + k, (int)c);
            }
        }
        // the claim counters go back to zero for the next frame (every CTA read them before barrier 3)
        for (int p = p0 + tid; p < p1; p += T) PA.ptotal[p] = 0u;
    }
    FUSED_TRACE(8);
    fused_grid_barrier(&A.ctrl->grid_bar, 4u * G);
    FUSED_TRACE(9);

    // ---- phase 4: global ranks, inverse, records -----------------------------------------------------------
    unsigned long long n_voxels = 0ull;
    if (ok) {
        // s_hist <- exclusive scan of the partition voxel counts
        for (int p = tid; p < nparts; p += T) s_hist[p] = __ldcg(PA.pvox + p);
        __syncthreads();
        n_voxels = cta_exclusive_scan_u32(s_hist, nparts, s_wsum, T);
        if (tid == 0) D.n_voxels = (int64_t)n_voxels;
        uint2* ring = s_ring + (size_t)warp * kFusedRing;
        unsigned head = 0, count = 0;
        auto drain = [&](unsigned slot, bool active) {
            if (!active) return;
            const uint2 e = ring[slot];
            const float4 q = s_pts[e.x];
            accumulate_dup(q, (int)s_key[e.x], e.y, D, A.acc, A.cnt);
        };
        const int m_round = (m + 31) & ~31;
        const int rot = m ? (int)(((long long)b * m / G) & ~31ll) : 0;      // as in the scatter: CTAs start at different partitions
        for (int t0 = tid; t0 < m_round; t0 += kFusedBatch * T) {
            int j[kFusedBatch], tt[kFusedBatch];
            unsigned ans[kFusedBatch], part[kFusedBatch];
            int key[kFusedBatch];
            float4 q[kFusedBatch];
            bool live[kFusedBatch];
#pragma unroll
            for (int u = 0; u < kFusedBatch; ++u) {
                const int i = t0 + u * T;
                live[u] = i < m;
                tt[u] = i + rot < m ? i + rot : i + rot - m;             // sorted position handled by this lane
                j[u] = live[u] ? (int)s_perm[tt[u]] : 0;
                q[u] = s_pts[j[u]];
            }
#pragma unroll
            for (int u = 0; u < kFusedBatch; ++u) {
                key[u] = (int)s_key[j[u]];
                part[u] = (unsigned)key[u] >> kPartShift;
                ans[u] = 0u;
                if (live[u]) ans[u] = __ldcg(PA.prank + s_delta[part[u]] + (unsigned)tt[u]);
            }
#pragma unroll
            for (int u = 0; u < kFusedBatch; ++u) {
                const unsigned r = s_hist[part[u]] + (ans[u] & ~kPartDup);
                const bool dup = live[u] && (ans[u] & kPartDup) != 0u;
                if (live[u]) {
                    s_inv[j[u]] = r;
                    if (!dup) store_first_member(A.voxels, r, q[u], key[u]);
                    if (K.do_grid && !cell16) {
                        const int cell = grid_cell_of(q[u], D, K);
                        if (cell >= 0) red_add_s32(A.grid_out + cell, 1);
                    }
                }
                const unsigned mm = __ballot_sync(0xffffffffu, dup);
                if (mm) {
                    if (dup) ring[(head + count + __popc(mm & lanemask_lt())) % kFusedRing] = make_uint2((unsigned)j[u], r);
                    count += __popc(mm);
                    __syncwarp();
                    if (count >= 32) {
                        drain((head + lane) % kFusedRing, true);
                        head = (head + 32) % kFusedRing;
                        count -= 32;
                        __syncwarp();
                    }
                }
            }
        }
        if (count) drain((head + lane) % kFusedRing, lane < count);
        __syncthreads();
        for (int jj = tid; jj < m; jj += T) st_stream_s32(A.inverse + c0 + jj, (int)s_inv[jj]);
    }
    FUSED_TRACE(10);
    fused_grid_barrier(&A.ctrl->grid_bar, 5u * G);
    FUSED_TRACE(11);

    // ---- phase 5: finalize the multi-member voxels ----------------------------------------------------------
    if (ok) {
        unsigned* ring = reinterpret_cast<unsigned*>(s_ring + (size_t)warp * kFusedRing);
        unsigned head = 0, count = 0;
        const double isx = 1.0 / D.fix_scale_xyz, isw = 1.0 / D.fix_scale_w;
        const int64_t V = (int64_t)n_voxels;
        const int64_t vper = ((V + G - 1) / G + 31) & ~(int64_t)31;
        const int64_t v0 = (int64_t)b * vper;
        const int64_t v1 = v0 + vper;
        for (int64_t r0 = v0 + tid; r0 < v1; r0 += (int64_t)kFusedBatch * T) {
            int c[kFusedBatch];
#pragma unroll
            for (int u = 0; u < kFusedBatch; ++u) {
                const int64_t r = r0 + (int64_t)u * T;
                c[u] = (r < v1 && r < V) ? __ldcg(A.cnt + r) : 0;
            }
#pragma unroll
            for (int u = 0; u < kFusedBatch; ++u) {
                const int64_t r = r0 + (int64_t)u * T;
                const bool hit = c[u] != 0;
                const unsigned mm = __ballot_sync(0xffffffffu, hit);
                if (mm) {
                    if (hit) ring[(head + count + __popc(mm & lanemask_lt())) % kFusedRing] = (unsigned)r;
                    count += __popc(mm);
                    __syncwarp();
                    if (count >= 32) {
                        finalize_voxel(ring[(head + lane) % kFusedRing], D, isx, isw, A.acc, A.cnt, A.voxels);
                        head = (head + 32) % kFusedRing;
                        count -= 32;
                        __syncwarp();
                    }
                }
            }
        }
        if (lane < count) finalize_voxel(ring[(head + lane) % kFusedRing], D, isx, isw, A.acc, A.cnt, A.voxels);
    }
    __syncthreads();
    if (tid == 0) {
        s_trace[12] = global_timer_ns();
        for (int k = 0; k < 13; ++k) A.trace_all[(size_t)b * 16 + k] = s_trace[k];
        if (b == 0) {
#pragma unroll
            for (int k = 0; k < 12; ++k) D.trace_ns[k] = (uint32_t)(s_trace[k + 1] - s_trace[k]);
            D.trace_ns[13] = 1u;                    // partitioned back end
            D.trace_ns[14] = 0u;
            D.trace_ns[15] = (uint32_t)G;
            *A.D = D;
        }
        __threadfence();
        if (atomicAdd(&A.ctrl->exit_ticket, 1u) == (unsigned)G - 1u) {
            A.ctrl->grid_bar = 0u;
            A.ctrl->grid_bar_b = 0u;
            A.ctrl->exit_ticket = 0u;
            A.ctrl->dirty_groups = 0ull;
        }
    }
}

// ---- k_frame_pack -----------------------------------------------------------------------------
// Host-bound results leave the device as structure-of-arrays: 16 B centroid + 4 B count (+ 4 B key) per
// voxel instead of the 32-byte record, i.e. 37 % fewer bytes over PCIe for the same information.
__global__ void __launch_bounds__(kFrameThreads)
k_frame_pack(const lidar_voxel* __restrict__ voxels, const lidar_frame_desc* __restrict__ Dg, int64_t cap,
             float4* __restrict__ centroids, int32_t* __restrict__ counts, int32_t* __restrict__ keys) {
    if (Dg->status != 0) return;
    int64_t V = Dg->n_voxels;
    if (V > cap) V = cap;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < V; r += stride) {
        const Word8 rec = ld_cg_256(voxels + r);
        centroids[r] = make_float4(__uint_as_float(rec.w[0]), __uint_as_float(rec.w[1]), __uint_as_float(rec.w[2]),
                                   __uint_as_float(rec.w[3]));
        counts[r] = (int32_t)rec.w[4];
        if (keys) keys[r] = (int32_t)rec.w[5];
    }
}

static int g_ctas_per_sm = 8;   // grid cap of the per-point frame kernels, in CTAs per SM (tuning knob)

// fused-kernel configuration (lidar_frame_set_fused)
static int g_fused_mode = LIDAR_FRAME_AUTO;
static int g_fused_threads = 512;
static int g_fused_ctas_per_sm = 1;
static int g_fused_smem_kb = 0;          // 0 = as much as the chunk needs, up to the opt-in maximum
static int g_fused_plain_launch = 0;     // experiment: ordinary launch instead of cooperative (see header)
static int g_fused_pdl = 0;              // programmatic dependent launch: frame f+1 loads while frame f drains
static thread_local int g_fused_scan_order = 0;   // 1: the scan-order variant of k_frame_fused (see the kernel's header).
                                                  // Per host thread: pipelines of different threads pick their variant independently
static size_t g_fused_l2_persist = 0;     // bytes of the occupancy groups pinned in L2 (access policy window), 0 = off
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) applies to the CURRENT device only: one flag per device
static bool g_fused_attr_set[64] = {};
static bool g_part_attr_set[64] = {};
static int g_part_auto = 0;              // LIDAR_FRAME_AUTO prefers the partitioned back end when a frame is eligible

static int frame_grid(int64_t n, int per_thread) {
    int64_t want = (n + (int64_t)kFrameThreads * per_thread - 1) / ((int64_t)kFrameThreads * per_thread);
    if (want < 1) want = 1;
    const int64_t cap = (int64_t)sm_count() * g_ctas_per_sm;
    return (int)(want < cap ? want : cap);
}

}  // namespace lidar

using namespace lidar;

extern "C" {

int lidar_frame_set_fused(int mode, int threads, int ctas_per_sm, int smem_kb) {
    LIDAR_REQUIRE(mode == LIDAR_FRAME_AUTO || mode == LIDAR_FRAME_MULTIKERNEL || mode == LIDAR_FRAME_FUSED ||
                      mode == LIDAR_FRAME_PARTITIONED,
                  LIDAR_ERR_INVALID, "lidar_frame_set_fused: unknown mode %d", mode);
    LIDAR_REQUIRE(threads == 0 || (threads >= 128 && threads <= 1024 && threads % 32 == 0),
                  LIDAR_ERR_INVALID, "lidar_frame_set_fused: threads must be a multiple of 32 in [128, 1024]");
    LIDAR_REQUIRE(ctas_per_sm >= 0 && ctas_per_sm <= 4, LIDAR_ERR_INVALID, "lidar_frame_set_fused: ctas_per_sm 0..4");
    LIDAR_REQUIRE(smem_kb >= 0 && smem_kb <= 227, LIDAR_ERR_INVALID, "lidar_frame_set_fused: smem_kb 0..227");
    g_fused_mode = mode;
    if (threads) g_fused_threads = threads;
    if (ctas_per_sm) g_fused_ctas_per_sm = ctas_per_sm;
    g_fused_smem_kb = smem_kb;
    return LIDAR_OK;
}

size_t lidar_frame_trace_offset(const lidar_frame_caps* caps) {
    if (!caps || caps->max_points < 0 || caps->max_key_space <= 0) return 0;
    return frame_layout(*caps).off_trace;
}

int lidar_frame_set_fused_plain_launch(int on) {
    g_fused_plain_launch = on ? 1 : 0;
    return LIDAR_OK;
}

int lidar_frame_set_fused_scan_order(int on) {
    g_fused_scan_order = on ? 1 : 0;
    return LIDAR_OK;
}

int lidar_frame_set_fused_l2_persist(size_t bytes) {
    if (bytes) {
        int dev = 0, max_persist = 0, max_window = 0;
        LIDAR_CUDA_TRY(cudaGetDevice(&dev));
        LIDAR_CUDA_TRY(cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev));
        LIDAR_CUDA_TRY(cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev));
        if (bytes > (size_t)max_persist) bytes = (size_t)max_persist;
        if (bytes > (size_t)max_window) bytes = (size_t)max_window;
        LIDAR_REQUIRE(bytes > 0, LIDAR_ERR_INVALID, "lidar_frame_set_fused_l2_persist: the device has no persisting L2");
        LIDAR_CUDA_TRY(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, bytes));
    }
    g_fused_l2_persist = bytes;
    return LIDAR_OK;
}

int lidar_frame_set_partition_auto(int on) {
    g_part_auto = on ? 1 : 0;
    return LIDAR_OK;
}

int lidar_frame_set_fused_pdl(int on) {
    LIDAR_REQUIRE(on >= 0 && on <= 2, LIDAR_ERR_INVALID, "lidar_frame_set_fused_pdl: 0 (off), 1 (on) or 2 (on, inputs complete)");
    g_fused_pdl = on;
    return LIDAR_OK;
}

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
    if (n > 0 && g_fused_mode != LIDAR_FRAME_MULTIKERNEL) {
        // ---- the whole frame as one persistent cooperative kernel -----------------------------
        const int T = g_fused_threads;
        const int G = sm_count() * g_fused_ctas_per_sm;
        LIDAR_REQUIRE(G <= kFusedMaxCtas, LIDAR_ERR_CAPACITY, "lidar_frame_voxel_density: fused grid too large");
        const size_t ring_bytes = (size_t)(T / 32) * kFusedRing * sizeof(uint2);
        const size_t static_bytes = 8192;    // static __shared__ of k_frame_fused (scan-order variant: + 4 KB of range bases), rounded up
        size_t budget = g_fused_smem_kb ? (size_t)g_fused_smem_kb * 1024 : smem_optin();
        if (budget > smem_optin()) budget = smem_optin();
        LIDAR_REQUIRE(budget > static_bytes + ring_bytes + 4096, LIDAR_ERR_INVALID,
                      "lidar_frame_voxel_density: fused shared-memory budget too small");
        budget -= static_bytes;
        const int64_t per = (((n + G - 1) / G) + 31) & ~(int64_t)31;
        const int64_t gpc = (L.groups + G - 1) / G;     // worst case (capacity), so the layout is per pipeline
        // group popcount cache first (1 B per group, capped at 16 KB), the rest holds points (20 B each)
        size_t gbytes = (size_t)((gpc + 127) & ~(int64_t)127);
        if (gbytes > 16384) gbytes = 0;
        int64_t room = ((int64_t)budget - (int64_t)ring_bytes - (int64_t)gbytes) / 20;
        room &= ~(int64_t)31;
        int64_t spts = per < room ? per : room;
        if (spts < 0) spts = 0;
        const size_t dyn = (size_t)spts * 20 + ring_bytes + gbytes;
        int cur_dev = 0;
        LIDAR_CUDA_TRY(cudaGetDevice(&cur_dev));
        if (cur_dev < 0 || cur_dev >= 64 || !g_fused_attr_set[cur_dev]) {
            LIDAR_CUDA_TRY(cudaFuncSetAttribute(k_frame_fused<false, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (int)(smem_optin() - static_bytes)));
            LIDAR_CUDA_TRY(cudaFuncSetAttribute(k_frame_fused<true, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (int)(smem_optin() - static_bytes)));
            LIDAR_CUDA_TRY(cudaFuncSetAttribute(k_frame_fused<false, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (int)(smem_optin() - static_bytes)));
            LIDAR_CUDA_TRY(cudaFuncSetAttribute(k_frame_fused<true, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (int)(smem_optin() - static_bytes)));
            if (cur_dev >= 0 && cur_dev < 64) g_fused_attr_set[cur_dev] = true;
        }
        FusedArgs A;
        A.P = P;
        A.partial = partial;
        A.ctrl = ctrl;
        A.D = d_desc;
        A.voxel_key = d_voxel_key;
        A.inverse = d_inverse;
        A.voxels = d_voxels;
        A.grid_out = d_grid;
        A.grid_cap = grid_cap;
        A.groups = groups;
        A.groups_cap = L.groups;
        A.acc = acc;
        A.cnt = cnt;
        A.cta_desc = reinterpret_cast<unsigned long long*>(ws + L.off_cta_desc);
        A.trace_all = reinterpret_cast<unsigned long long*>(ws + L.off_trace);
        A.grid_rep = reinterpret_cast<int32_t*>(ws + L.off_grid_rep);
        A.smem_points = (int)spts;
        A.smem_groups = (int)gbytes;
        A.early_load = g_fused_pdl == 2 ? 1 : 0;
        A.l1 = reinterpret_cast<uint32_t*>(ws + L.off_l1);
        A.l1cnt = reinterpret_cast<uint32_t*>(ws + L.off_l1cnt);
        // ---- partitioned back end: eligible when every CTA's chunk is resident and the key space fits 2048 partitions
        bool use_part = false;
        PartArgs PA{};
        size_t dyn_part = 0;
        if (g_fused_mode == LIDAR_FRAME_PARTITIONED || (g_fused_mode == LIDAR_FRAME_AUTO && g_part_auto && !g_fused_scan_order)) {
            const int64_t pcap = ((((caps->max_key_space + kPartCells - 1) >> kPartShift) + 31) & ~(int64_t)31);
            const size_t ring_part = ring_bytes > 8192 ? ring_bytes : 8192;          // doubles as the partition-base array
            const size_t owner_bytes = (size_t)kPartWords64 * 12 + (size_t)kPartDensCap * 4;
            size_t union_bytes = (size_t)per * 4 > owner_bytes ? (size_t)per * 4 : owner_bytes;   // {cell, slot} | owner | inverse
            union_bytes = (union_bytes + 127) & ~(size_t)127;
            dyn_part = (size_t)per * 22 + (size_t)pcap * 8 + ring_part + union_bytes;
            const bool fits = pcap <= kPartMax && dyn_part + static_bytes <= smem_optin() && per < 65536 && (T & (T - 1)) == 0 &&
                              T <= kFusedMaxThreads;
            LIDAR_REQUIRE(fits || g_fused_mode != LIDAR_FRAME_PARTITIONED, LIDAR_ERR_CAPACITY,
                          "lidar_frame_voxel_density: the partitioned back end needs caps.max_key_space <= 2^29 and the frame "
                          "resident in shared memory (%lld points per CTA, %zu B of %zu B)", (long long)per,
                          dyn_part + static_bytes, smem_optin());
            if (fits) {
                use_part = true;
                PA.F = A;
                PA.F.smem_points = (int)per;
                PA.ptotal = reinterpret_cast<unsigned*>(ws + L.off_ptotal);
                PA.pvox = reinterpret_cast<unsigned*>(ws + L.off_pvox);
                PA.bucket = reinterpret_cast<uint2*>(ws + L.off_bucket);
                PA.prank = reinterpret_cast<unsigned*>(ws + L.off_prank);
                PA.pcap = (int)pcap;
                PA.union_bytes = (int)union_bytes;
                if (cur_dev < 0 || cur_dev >= 64 || !g_part_attr_set[cur_dev]) {
                    LIDAR_CUDA_TRY(cudaFuncSetAttribute(k_frame_part, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                        (int)(smem_optin() - static_bytes)));
                    if (cur_dev >= 0 && cur_dev < 64) g_part_attr_set[cur_dev] = true;
                }
            }
        }
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(G);
        cfg.blockDim = dim3(T);
        cfg.dynamicSmemBytes = use_part ? dyn_part : dyn;
        cfg.stream = st;
        cudaLaunchAttribute attr[3];
        int na = 0;
        if (g_fused_l2_persist) {
            // the occupancy groups are the one structure every point hits twice at random: keep them resident in L2
            // while the frames (read once) and the outputs (written once) stream past
            size_t bytes = (size_t)L.groups * 32;
            if (bytes > g_fused_l2_persist) bytes = g_fused_l2_persist;
            attr[na].id = cudaLaunchAttributeAccessPolicyWindow;
            attr[na].val.accessPolicyWindow.base_ptr = groups;
            attr[na].val.accessPolicyWindow.num_bytes = bytes;
            attr[na].val.accessPolicyWindow.hitRatio = 1.0f;
            attr[na].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            attr[na].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            ++na;
        }
        if (!g_fused_plain_launch) {
            attr[na].id = cudaLaunchAttributeCooperative;      // gang scheduling: the grid barriers cannot deadlock
            attr[na].val.cooperative = 1;
            ++na;
        }
        if (g_fused_pdl) {
            attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[na].val.programmaticStreamSerializationAllowed = 1;
            ++na;
        }
        cfg.attrs = attr;
        cfg.numAttrs = na;
        cudaError_t le = use_part ? cudaLaunchKernelEx(&cfg, k_frame_part, PA)
                         : T > 512 ? (g_fused_scan_order ? cudaLaunchKernelEx(&cfg, k_frame_fused<true, 1024>, A)
                                                         : cudaLaunchKernelEx(&cfg, k_frame_fused<false, 1024>, A))
                                   : (g_fused_scan_order ? cudaLaunchKernelEx(&cfg, k_frame_fused<true, 512>, A)
                                                         : cudaLaunchKernelEx(&cfg, k_frame_fused<false, 512>, A));
        if (le == cudaSuccess) {
            for (int i = 1; i <= 5; ++i) LIDAR_CUDA_TRY(mark(i));
            return LIDAR_OK;
        }
        (void)cudaGetLastError();
        LIDAR_REQUIRE(g_fused_mode == LIDAR_FRAME_AUTO, LIDAR_ERR_CUDA,
                      "lidar_frame_voxel_density: cooperative launch of k_frame_fused failed: %s",
                      cudaGetErrorString(le));
        // AUTO: fall through to the five-kernel path (e.g. the grid cannot be co-resident)
    }
    int bgrid = frame_grid(n > 0 ? n : 1, 4);
    if (bgrid < sm_count()) bgrid = sm_count();   // enough CTAs to zero the bitmap quickly
    if (bgrid > kBboxMaxBlocks) bgrid = kBboxMaxBlocks;
    k_frame_prep<<<bgrid, kFrameThreads, 0, st>>>(P, partial, ctrl, d_desc, d_grid, grid_cap, tile_desc, L.tiles,
                                                  groups, L.groups);
    LIDAR_CHECK_LAUNCH();
    LIDAR_CUDA_TRY(mark(1));
    if (n == 0) {
        for (int i = 2; i <= 5; ++i) LIDAR_CUDA_TRY(mark(i));
        return LIDAR_OK;
    }
    k_frame_mark<<<frame_grid(n, 2), kFrameThreads, 0, st>>>(P.pts, d_desc, d_voxel_key, groups, d_grid);
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
    k_frame_rank<<<frame_grid(n, 2), kFrameThreads, 0, st>>>(P.pts, d_desc, d_voxel_key, groups, d_inverse,
                                                              acc, cnt, d_voxels);
    LIDAR_CHECK_LAUNCH();
    LIDAR_CUDA_TRY(mark(4));
    k_frame_finalize<<<frame_grid(n, 4), kFrameThreads, 0, st>>>(d_desc, acc, cnt, d_voxels);
    LIDAR_CHECK_LAUNCH();
    LIDAR_CUDA_TRY(mark(5));
    return LIDAR_OK;
}

int lidar_frame_pack_soa(const lidar_voxel* d_voxels, const lidar_frame_desc* d_desc, int64_t capacity,
                         float* d_centroids4, int32_t* d_counts, int32_t* d_keys, void* stream) {
    LIDAR_REQUIRE(d_voxels && d_desc && d_centroids4 && d_counts, LIDAR_ERR_INVALID, "lidar_frame_pack_soa: NULL argument");
    LIDAR_REQUIRE(capacity >= 0, LIDAR_ERR_INVALID, "lidar_frame_pack_soa: negative capacity");
    if (capacity == 0) return LIDAR_OK;
    k_frame_pack<<<frame_grid(capacity, 4), kFrameThreads, 0, as_stream(stream)>>>(
        d_voxels, d_desc, capacity, reinterpret_cast<float4*>(d_centroids4), d_counts, d_keys);
    LIDAR_CHECK_LAUNCH();
    return LIDAR_OK;
}

// ---- host-buffer entry: one call = copy-in + frame + repack + ONE copy-out -------------------------
static void host_block_offsets(int64_t n, const lidar_frame_caps& c, int flags, size_t off[8]) {
    const size_t np = (size_t)((n + 7) & ~(int64_t)7);
    const size_t grid_cells = (size_t)(c.max_nx > 0 ? c.max_nx : 0) * (size_t)(c.max_ny > 0 ? c.max_ny : 0);
    size_t o = 0;
    auto take = [&](size_t bytes) { const size_t a = (o + 31) & ~(size_t)31; o = a + bytes; return a; };
    off[0] = take(4 * np);          // voxel_key
    off[1] = take(4 * np);          // inverse
    off[2] = take(16 * np);         // centroids
    off[3] = take(4 * np);          // counts
    off[4] = take((flags & LIDAR_HOST_UNIQUE_KEYS) ? 4 * np : 0);   // unique keys
    off[5] = take(4 * grid_cells);  // grid
    off[6] = take(sizeof(lidar_frame_desc));
    off[7] = (o + 31) & ~(size_t)31;
}

size_t lidar_frame_host_block_bytes(int64_t n, const lidar_frame_caps* caps, int flags) {
    if (!caps || n < 0) return 0;
    size_t off[8];
    host_block_offsets(n, *caps, flags, off);
    return off[7];
}

int lidar_frame_host_block_layout(int64_t n, const lidar_frame_caps* caps, int flags, size_t* h_offsets7) {
    LIDAR_REQUIRE(caps && h_offsets7 && n >= 0, LIDAR_ERR_INVALID, "lidar_frame_host_block_layout: bad argument");
    size_t off[8];
    host_block_offsets(n, *caps, flags, off);
    for (int i = 0; i < 7; ++i) h_offsets7[i] = off[i];
    return LIDAR_OK;
}

// copy-in + frame + repack; `copy_all`: the whole result block follows in one copy sized by the FRAME (the per-voxel
// arrays travel at capacity n because n_voxels is not known on the host yet); otherwise only the descriptor comes
// back and lidar_frame_host_fetch() copies exactly what the frame produced once the host has read it
static int frame_host_begin(const void* h_points, int64_t n, double voxel_size, double grid_size,
                            const double* h_origin3, const double* h_xy_range4, void* d_points,
                            lidar_voxel* d_voxels, void* d_out, void* h_out, int flags,
                            const lidar_frame_caps* caps, void* d_ws, size_t ws_bytes, void* stream, bool copy_all) {
    LIDAR_REQUIRE(caps != nullptr, LIDAR_ERR_INVALID, "lidar_frame_voxel_density_host: caps is NULL");
    LIDAR_REQUIRE(n >= 0 && n <= caps->max_points, LIDAR_ERR_CAPACITY,
                  "lidar_frame_voxel_density_host: n=%lld exceeds caps.max_points=%lld", (long long)n,
                  (long long)caps->max_points);
    LIDAR_REQUIRE(d_out && h_out && d_voxels && (n == 0 || (h_points && d_points)), LIDAR_ERR_INVALID,
                  "lidar_frame_voxel_density_host: NULL buffer");
    size_t off[8];
    host_block_offsets(n, *caps, flags, off);
    char* out = static_cast<char*>(d_out);
    cudaStream_t st = as_stream(stream);
    if (n > 0) LIDAR_CUDA_TRY(cudaMemcpyAsync(d_points, h_points, (size_t)n * 16, cudaMemcpyHostToDevice, st));
    lidar_frame_desc* d_desc = reinterpret_cast<lidar_frame_desc*>(out + off[6]);
    int32_t* d_grid = grid_size > 0.0 ? reinterpret_cast<int32_t*>(out + off[5]) : nullptr;
    const int rc = frame_voxel_density_impl(d_points, n, voxel_size, grid_size, h_origin3, h_xy_range4,
                                            reinterpret_cast<int32_t*>(out + off[0]), reinterpret_cast<int32_t*>(out + off[1]),
                                            d_voxels, d_grid, d_desc, caps, d_ws, ws_bytes, stream, nullptr);
    if (rc != LIDAR_OK) return rc;
    if (n > 0) {
        k_frame_pack<<<frame_grid(n, 4), kFrameThreads, 0, st>>>(d_voxels, d_desc, n, reinterpret_cast<float4*>(out + off[2]),
                                                                 reinterpret_cast<int32_t*>(out + off[3]),
                                                                 (flags & LIDAR_HOST_UNIQUE_KEYS) ? reinterpret_cast<int32_t*>(out + off[4]) : nullptr);
        LIDAR_CHECK_LAUNCH();
    }
    if (copy_all) {
        // one copy-out: the grid is zero-filled to its capacity by the frame, the per-voxel tails are never read;
        // without the per-point outputs the copy starts at the centroids
        const size_t first = (flags & LIDAR_HOST_NO_PER_POINT) ? off[2] : 0;
        LIDAR_CUDA_TRY(cudaMemcpyAsync(static_cast<char*>(h_out) + first, out + first, off[7] - first, cudaMemcpyDeviceToHost, st));
    } else {
        LIDAR_CUDA_TRY(cudaMemcpyAsync(static_cast<char*>(h_out) + off[6], out + off[6], sizeof(lidar_frame_desc),
                                       cudaMemcpyDeviceToHost, st));
    }
    return LIDAR_OK;
}

int lidar_frame_voxel_density_host(const void* h_points, int64_t n, double voxel_size, double grid_size,
                                   const double* h_origin3, const double* h_xy_range4, void* d_points,
                                   lidar_voxel* d_voxels, void* d_out, void* h_out, int flags,
                                   const lidar_frame_caps* caps, void* d_ws, size_t ws_bytes, void* stream) {
    return frame_host_begin(h_points, n, voxel_size, grid_size, h_origin3, h_xy_range4, d_points, d_voxels, d_out, h_out,
                            flags, caps, d_ws, ws_bytes, stream, true);
}

int lidar_frame_voxel_density_host_begin(const void* h_points, int64_t n, double voxel_size, double grid_size,
                                         const double* h_origin3, const double* h_xy_range4, void* d_points,
                                         lidar_voxel* d_voxels, void* d_out, void* h_out, int flags,
                                         const lidar_frame_caps* caps, void* d_ws, size_t ws_bytes, void* stream) {
    return frame_host_begin(h_points, n, voxel_size, grid_size, h_origin3, h_xy_range4, d_points, d_voxels, d_out, h_out,
                            flags, caps, d_ws, ws_bytes, stream, false);
}

int lidar_frame_host_fetch(int64_t n, int64_t n_voxels, int nx, int ny, const void* d_out, void* h_out, int flags,
                           const lidar_frame_caps* caps, void* stream) {
    LIDAR_REQUIRE(caps && d_out && h_out && n >= 0 && n_voxels >= 0 && n_voxels <= n && nx >= 0 && ny >= 0,
                  LIDAR_ERR_INVALID, "lidar_frame_host_fetch: bad argument");
    LIDAR_REQUIRE((int64_t)nx * ny <= (int64_t)(caps->max_nx > 0 ? caps->max_nx : 0) * (caps->max_ny > 0 ? caps->max_ny : 0),
                  LIDAR_ERR_CAPACITY, "lidar_frame_host_fetch: grid %dx%d exceeds the capacities", nx, ny);
    size_t off[8];
    host_block_offsets(n, *caps, flags, off);
    const char* src = static_cast<const char*>(d_out);
    char* dst = static_cast<char*>(h_out);
    cudaStream_t st = as_stream(stream);
    auto copy = [&](size_t o, size_t bytes) -> cudaError_t {
        return bytes ? cudaMemcpyAsync(dst + o, src + o, bytes, cudaMemcpyDeviceToHost, st) : cudaSuccess;
    };
    if (!(flags & LIDAR_HOST_NO_PER_POINT)) LIDAR_CUDA_TRY(copy(off[0], off[1] - off[0] + (size_t)n * 4));   // key | inverse
    LIDAR_CUDA_TRY(copy(off[2], (size_t)n_voxels * 16));
    LIDAR_CUDA_TRY(copy(off[3], (size_t)n_voxels * 4));
    if (flags & LIDAR_HOST_UNIQUE_KEYS) LIDAR_CUDA_TRY(copy(off[4], (size_t)n_voxels * 4));
    LIDAR_CUDA_TRY(copy(off[5], (size_t)nx * ny * 4));
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
