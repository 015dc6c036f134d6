// Point-sharded density grid of ONE oversized scan (BASELINE configs[4]; SURVEY.md §8e "points").
//
// calculate_grid_density (utils/data_processing.py:282-328) of a scan whose points are spread over the GPUs
// of one box: every rank bins its shard into the SAME np.arange edges and the integer grids are summed
// (order independent => bit-identical to the single-GPU grid).
//
// Two forms, same arithmetic:
//
//  (1) lidar_scan_density — ONE persistent cooperative kernel per rank does the whole call:
//        local bbox -> [bbox exchange with the peers: plain stores into their symmetric buffers + flags]
//        -> arange parameters derived ON THE DEVICE (edges.cuh) -> histogram (RED.ADD into the rank's grid)
//        -> [grid all-reduce over NVLink, two-shot: each rank reduces ITS slice of the grid across all
//            replicas with multimem.ld_reduce.add.u32 (NVSwitch reduces in the switch, NVLS) and broadcasts
//            it with multimem.st; peer-pointer loads / stores when the buffer has no multicast mapping]
//        -> density = counts / g^2 and the cell centres in fp64.
//      The descriptor (nx, ny, status) is also written to MAPPED HOST memory as soon as the edges are
//      known, so the host can enqueue the exactly-sized read-back behind the kernel without waiting for it.
//      The symmetric buffer (flags | bbox slots | grid) is allocated by the caller (torch symmetric
//      memory is the plumbing: it hands out the peer and multicast addresses); world = 1 needs none of it.
//
//  (2) lidar_scan_bbox_packed / lidar_scan_hist / lidar_scan_finish — the same phases as three enqueues for
//      callers that bring their own collectives (lidar_nccl_* in nccl.cu, or torch.distributed): MAX over
//      the packed bbox, SUM over the int32 grid.  Nothing visits the host in between.
#include <cooperative_groups.h>

#include "common.cuh"
#include "edges.cuh"

namespace lidar {

constexpr int kScanThreads = 256;
constexpr int kScanMaxCtas = 2048;
constexpr int kScanMaxWorld = 16;

// layout of the symmetric buffer (same on every rank); the grid follows at kSymmGridOffset
struct ScanSymmHeader {
    unsigned flags[4][kScanMaxWorld];     // [phase][source rank]: epoch counters written by the peers
    double bbox[kScanMaxWorld][4];        // [source rank]{-minx, -miny, maxx, maxy}
};
constexpr size_t kSymmGridOffset = 1024;
static_assert(sizeof(ScanSymmHeader) <= kSymmGridOffset, "symmetric header grew past the grid offset");

struct ScanWs {
    unsigned bar;                 // grid barrier arrivals (reset by the last CTA to leave)
    unsigned exit_ticket;
    unsigned pad[2];
    lidar_scan_desc desc;         // device copy of the descriptor (every CTA reads it after barrier 2)
};                                // followed by double partial[kScanMaxCtas][4]: per-CTA {-minx, -miny, maxx, maxy}

struct ScanArgs {
    const void* pts;
    int fmt;
    int64_t n;
    double g;
    int max_nx, max_ny;
    int64_t cap_cells;
    ScanWs* ws;
    int32_t* grid;                // this rank's grid (inside the symmetric buffer when world > 1)
    double* density;              // [nx*ny] out
    double* gx;                   // [nx] out
    double* gy;                   // [ny] out
    lidar_scan_desc* desc_out;    // device descriptor out
    lidar_scan_desc* desc_host;   // mapped host descriptor (may be NULL)
    // multi-GPU
    int rank, world;
    unsigned epoch;               // > 0, increases by one per call on every rank
    ScanSymmHeader* self;         // this rank's symmetric buffer
    ScanSymmHeader* peers[kScanMaxWorld];   // every rank's buffer as mapped on this device (peers[rank] == self)
    char* mc;                     // multicast mapping of the buffer (NULL: peer loads / stores)
};

__device__ __forceinline__ unsigned ld_acquire_sys_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_u32(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_gpu_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_gpu_add_u32(unsigned* p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned multimem_ld_reduce_add_u32(const void* mc) {
    unsigned v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.u32 %0, [%1];" : "=r"(v) : "l"(mc) : "memory");
    return v;
}
__device__ __forceinline__ void multimem_st_u32(void* mc, unsigned v) {
    asm volatile("multimem.st.relaxed.sys.global.u32 [%0], %1;" ::"l"(mc), "r"(v) : "memory");
}
// Two adjacent 32-bit counters as ONE 64-bit add: the sum of (hi << 32 | lo) over the ranks is (sum hi) << 32 | (sum lo)
// as long as sum lo < 2^32 -- a density cell counts points, and a scan has fewer than 2^31 of them -- so the low word
// never carries into the high one and a single switch operation reduces two cells.
__device__ __forceinline__ unsigned long long multimem_ld_reduce_add_u64(const void* mc) {
    unsigned long long v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.u64 %0, [%1];" : "=l"(v) : "l"(mc) : "memory");
    return v;
}
__device__ __forceinline__ void multimem_st_u64(void* mc, unsigned long long v) {
    asm volatile("multimem.st.relaxed.sys.global.u64 [%0], %1;" ::"l"(mc), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long scan_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// counter barrier over the co-resident grid (cooperative launch)
__device__ __forceinline__ void scan_grid_barrier(unsigned* ctr, unsigned target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        red_release_gpu_add_u32(ctr, 1u);
        unsigned polls = 0;
        unsigned long long t0 = 0ull;
        while (ld_acquire_gpu_u32(ctr) < target) {
            if ((++polls & 0x3ffu) == 0u) {
                const unsigned long long t = scan_timer_ns();
                if (t0 == 0ull) t0 = t;
                else if (t - t0 > 4000000000ull) __trap();   // a CTA died: fail the launch instead of hanging
            }
        }
    }
    __syncthreads();
}

// Cross-rank barrier, called by the threads of ONE CTA after a grid barrier: lane r tells rank r "I reached
// `phase` of call `epoch`" and waits for rank r to say the same.  The release at system scope publishes
// everything this GPU wrote before the grid barrier (the barrier's acquire ordered it before this thread).
__device__ __forceinline__ void scan_rank_barrier(const ScanArgs& A, int phase) {
    if ((int)threadIdx.x < A.world) {
        const int r = threadIdx.x;
        __threadfence_system();
        st_release_sys_u32(&A.peers[r]->flags[phase][A.rank], A.epoch);
        unsigned polls = 0;
        unsigned long long t0 = 0ull;
        while ((int)(ld_acquire_sys_u32(&A.self->flags[phase][r]) - A.epoch) < 0) {
            if ((++polls & 0xffu) == 0u) {
                const unsigned long long t = scan_timer_ns();
                if (t0 == 0ull) t0 = t;
                else if (t - t0 > 10000000000ull) __trap();  // a peer never arrived (10 s)
            }
        }
    }
    __syncthreads();
}

template <class Loader>
__device__ __forceinline__ void scan_fold_xy(const Loader& L, int64_t i, float mn[2], float mx[2]);

template <>
__device__ __forceinline__ void scan_fold_xy<LoadF32x4>(const LoadF32x4& L, int64_t i, float mn[2], float mx[2]) {
    const float4 v = L.raw(i);
    mn[0] = fminf(mn[0], v.x); mx[0] = fmaxf(mx[0], v.x);
    mn[1] = fminf(mn[1], v.y); mx[1] = fmaxf(mx[1], v.y);
}

// bbox partial of this CTA over a grid-stride walk (x,y only: the density grid is 2-D).  fp32 min/max of fp32
// inputs is exact; the F64X3 layout folds in fp64 (DLoader below).
template <class Loader>
struct BboxFold;

template <>
struct BboxFold<LoadF32x4> {
    float mn[2] = {INFINITY, INFINITY}, mx[2] = {-INFINITY, -INFINITY};
    __device__ __forceinline__ void add(const LoadF32x4& L, int64_t i) { scan_fold_xy(L, i, mn, mx); }
    __device__ __forceinline__ double lo(int c) const { return (double)mn[c]; }
    __device__ __forceinline__ double hi(int c) const { return (double)mx[c]; }
};
template <>
struct BboxFold<LoadF64x3> {
    double mn[2] = {INFINITY, INFINITY}, mx[2] = {-INFINITY, -INFINITY};
    __device__ __forceinline__ void add(const LoadF64x3& L, int64_t i) {
        const Pt p = L.load(i);
        mn[0] = fmin(mn[0], p.x); mx[0] = fmax(mx[0], p.x);
        mn[1] = fmin(mn[1], p.y); mx[1] = fmax(mx[1], p.y);
    }
    __device__ __forceinline__ double lo(int c) const { return mn[c]; }
    __device__ __forceinline__ double hi(int c) const { return mx[c]; }
};

// CTA-wide fold of {-minx, -miny, maxx, maxy}: result valid in thread 0..3 of the CTA (s_out[4])
__device__ __forceinline__ void cta_max4(double v[4], double* s_red /* [warps][4] */, double* s_out) {
    const int warp = threadIdx.x >> 5;
#pragma unroll
    for (int c = 0; c < 4; ++c) v[c] = warp_max(v[c]);
    if (lane_id() == 0) {
#pragma unroll
        for (int c = 0; c < 4; ++c) s_red[warp * 4 + c] = v[c];
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double r = -INFINITY;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) r = fmax(r, s_red[w * 4 + threadIdx.x]);
        s_out[threadIdx.x] = r;
    }
    __syncthreads();
}

// descriptor from the global packed bbox {-minx, -miny, maxx, maxy}
__device__ __forceinline__ void scan_derive_desc(const double packed[4], double g, int max_nx, int max_ny,
                                                 int64_t cap_cells, int64_t n_local, lidar_scan_desc* D) {
    const double minx = -packed[0], miny = -packed[1], maxx = packed[2], maxy = packed[3];
    D->bbox[0] = minx; D->bbox[1] = miny; D->bbox[2] = maxx; D->bbox[3] = maxy;
    D->grid = g;
    D->n_local = n_local;
    // an empty scan keeps +inf / -inf: calculate_grid_density returns (None, None, None) (data_processing.py:297-298)
    const bool have = minx <= maxx && miny <= maxy && isfinite(minx) && isfinite(maxx) && isfinite(miny) && isfinite(maxy);
    int status = have ? 0 : LIDAR_SCAN_EMPTY;
    ArangeAxis X{}, Y{};
    if (have) {
        X = arange_axis(minx, maxx, g, max_nx, true);
        Y = arange_axis(miny, maxy, g, max_ny, true);
        status = X.status ? X.status : Y.status;
        if (!status && (int64_t)X.nb * Y.nb > cap_cells) status = LIDAR_ERR_CAPACITY;
    }
    D->ex0 = X.a; D->ex1 = X.e1; D->exd = X.delta; D->nx = status ? 0 : X.nb;
    D->ey0 = Y.a; D->ey1 = Y.e1; D->eyd = Y.delta; D->ny = status ? 0 : Y.nb;
    D->status = status;
    D->pad = 0;
}

struct ScanBinConst {
    float axf, ayf, rdxf, rdyf;
    double rdx, rdy;
};
__device__ __forceinline__ ScanBinConst scan_bin_const(const lidar_scan_desc& D) {
    ScanBinConst K;
    K.rdx = __ddiv_rn(1.0, D.exd);
    K.rdy = __ddiv_rn(1.0, D.eyd);
    K.axf = (float)D.ex0; K.ayf = (float)D.ey0; K.rdxf = (float)K.rdx; K.rdyf = (float)K.rdy;
    return K;
}
__device__ __forceinline__ void scan_red_add(int32_t* p, int v) {
    asm volatile("red.relaxed.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <class Loader>
__device__ __forceinline__ int scan_cell(const Loader& L, int64_t i, const lidar_scan_desc& D, const ScanBinConst& K);
template <>
__device__ __forceinline__ int scan_cell<LoadF32x4>(const LoadF32x4& L, int64_t i, const lidar_scan_desc& D, const ScanBinConst& K) {
    const float4 q = L.raw(i);
    const int bx = fast_arange_bin(q.x, (double)q.x, K.axf, K.rdxf, D.ex0, D.ex1, D.exd, K.rdx, D.nx);
    const int by = fast_arange_bin(q.y, (double)q.y, K.ayf, K.rdyf, D.ey0, D.ey1, D.eyd, K.rdy, D.ny);
    return (bx >= 0 && by >= 0) ? bx * D.ny + by : -1;
}
template <>
__device__ __forceinline__ int scan_cell<LoadF64x3>(const LoadF64x3& L, int64_t i, const lidar_scan_desc& D, const ScanBinConst& K) {
    const Pt p = L.load(i);
    const int bx = arange_bin(p.x, D.ex0, D.ex1, D.exd, K.rdx, D.nx);
    const int by = arange_bin(p.y, D.ey0, D.ey1, D.eyd, K.rdy, D.ny);
    return (bx >= 0 && by >= 0) ? bx * D.ny + by : -1;
}

// histogram of the shard into `grid` (four loads and four fire-and-forget REDs in flight per thread)
template <class Loader>
__device__ __forceinline__ void scan_hist_phase(const Loader& L, int64_t n, const lidar_scan_desc& D, int32_t* grid) {
    const ScanBinConst K = scan_bin_const(D);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n; i += 4 * stride) {
        int c[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) c[u] = scan_cell(L, i + u * stride, D, K);
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (c[u] >= 0) scan_red_add(grid + c[u], 1);
    }
    for (; i < n; i += stride) {
        const int c = scan_cell(L, i, D, K);
        if (c >= 0) scan_red_add(grid + c, 1);
    }
}

// density = counts / g^2 and the cell centres (data_processing.py:322-326), all fp64 IEEE operations
__device__ __forceinline__ void scan_finish_phase(const lidar_scan_desc& D, const int32_t* grid, double* density,
                                                  double* gx, double* gy) {
    if (D.status != 0) return;
    const double g2 = __dmul_rn(D.grid, D.grid);
    const int64_t cells = (int64_t)D.nx * D.ny;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int64_t k = t0; k < cells; k += stride) density[k] = __ddiv_rn((double)__ldcg(grid + k), g2);
    for (int64_t k = t0; k < D.nx; k += stride)
        gx[k] = __dmul_rn(__dadd_rn(arange_edge(D.ex0, D.ex1, D.exd, (int)k), arange_edge(D.ex0, D.ex1, D.exd, (int)k + 1)), 0.5);
    for (int64_t k = t0; k < D.ny; k += stride)
        gy[k] = __dmul_rn(__dadd_rn(arange_edge(D.ey0, D.ey1, D.eyd, (int)k), arange_edge(D.ey0, D.ey1, D.eyd, (int)k + 1)), 0.5);
}

// ================================================================================================
// (1) the whole call as one persistent cooperative kernel
// ================================================================================================
template <class Loader>
__global__ void __launch_bounds__(kScanThreads)
k_scan_density(const ScanArgs A) {
    __shared__ double s_red[kScanThreads / 32 * 4];
    __shared__ double s_bb[4];
    __shared__ lidar_scan_desc D;
    const Loader L{static_cast<decltype(Loader::p)>(A.pts)};
    const int G = gridDim.x, b = blockIdx.x, tid = threadIdx.x;
    const int64_t stride = (int64_t)G * blockDim.x;
    const int64_t t0 = (int64_t)b * blockDim.x + tid;
    unsigned bar_no = 0;

    // ---- phase A: zero this rank's grid, bbox partial of the shard --------------------------------
    {
        int4* g4 = reinterpret_cast<int4*>(A.grid);
        const int4 z = make_int4(0, 0, 0, 0);
        for (int64_t k = t0; k < A.cap_cells / 4; k += stride) g4[k] = z;
        BboxFold<Loader> F;
        int64_t i = t0;
        for (; i + 3 * stride < A.n; i += 4 * stride) { F.add(L, i); F.add(L, i + stride); F.add(L, i + 2 * stride); F.add(L, i + 3 * stride); }
        for (; i < A.n; i += stride) F.add(L, i);
        double v[4] = {-F.lo(0), -F.lo(1), F.hi(0), F.hi(1)};
        cta_max4(v, s_red, s_bb);
    }
    double* dpart = reinterpret_cast<double*>(reinterpret_cast<char*>(A.ws) + sizeof(ScanWs));
    if (tid < 4) dpart[(size_t)b * 4 + tid] = s_bb[tid];
    scan_grid_barrier(&A.ws->bar, ++bar_no * G);

    // ---- CTA 0: fold the partials, exchange with the peers, derive the descriptor ------------------
    if (b == 0) {
        double v[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        for (int q = tid; q < G; q += blockDim.x) {
#pragma unroll
            for (int c = 0; c < 4; ++c) v[c] = fmax(v[c], __ldcg(dpart + (size_t)q * 4 + c));
        }
        cta_max4(v, s_red, s_bb);
        if (A.world > 1) {
            // my packed bbox into every peer's slot [rank], then the flag exchange, then the MAX over ranks
            if (tid < A.world * 4) {
                const int r = tid >> 2, c = tid & 3;
                A.peers[r]->bbox[A.rank][c] = s_bb[c];
            }
            __syncthreads();
            scan_rank_barrier(A, 0);
            if (tid < 4) {
                double r = -INFINITY;
                for (int q = 0; q < A.world; ++q) r = fmax(r, __ldcg(&A.self->bbox[q][tid]));
                s_bb[tid] = r;
            }
            __syncthreads();
        }
        if (tid == 0) {
            scan_derive_desc(s_bb, A.g, A.max_nx, A.max_ny, A.cap_cells, A.n, &D);
            A.ws->desc = D;
            *A.desc_out = D;
            if (A.desc_host) {
                // mapped host memory: the host polls `epoch` and enqueues the exactly-sized read-back behind this kernel
                // fields first, then (release, system scope) the epoch: the host never sees a half-written descriptor
                lidar_scan_desc H = D;
                H.pad = 0;
                *A.desc_host = H;
                __threadfence_system();
                st_release_sys_u32(reinterpret_cast<unsigned*>(&A.desc_host->pad), A.epoch);
            }
        }
    }
    scan_grid_barrier(&A.ws->bar, ++bar_no * G);
    if (tid == 0 && b != 0) {
        const lidar_scan_desc* src = &A.ws->desc;
        lidar_scan_desc tmp;
        const unsigned long long* s8 = reinterpret_cast<const unsigned long long*>(src);
        unsigned long long* d8 = reinterpret_cast<unsigned long long*>(&tmp);
        for (size_t k = 0; k < sizeof(lidar_scan_desc) / 8; ++k) d8[k] = __ldcg(s8 + k);
        D = tmp;
    }
    __syncthreads();

    // ---- phase B: histogram ------------------------------------------------------------------------
    const bool ok = D.status == 0;
    if (ok) scan_hist_phase(L, A.n, D, A.grid);

    // ---- phase C: grid all-reduce over NVLink (two-shot) ------------------------------------------
    if (A.world > 1) {
        scan_grid_barrier(&A.ws->bar, ++bar_no * G);
        if (b == 0) scan_rank_barrier(A, 1);             // every rank's grid is complete
        scan_grid_barrier(&A.ws->bar, ++bar_no * G);
        if (ok) {
            // slices in PAIRS of cells (8-byte units); the grid is padded to a multiple of 4 cells and zero beyond nx*ny
            const int64_t pairs = ((int64_t)D.nx * D.ny + 1) / 2;
            const int64_t per = (pairs + A.world - 1) / A.world;
            const int64_t s0 = (int64_t)A.rank * per < pairs ? (int64_t)A.rank * per : pairs;
            const int64_t s1 = s0 + per < pairs ? s0 + per : pairs;
            if (A.mc) {
                // NVLS: the switch adds the replicas of a word and hands back the sum; the sum goes to every replica
                char* mcg = A.mc + kSymmGridOffset;
                for (int64_t k = s0 + t0; k < s1; k += stride) {
                    const unsigned long long v = multimem_ld_reduce_add_u64(mcg + 8 * k);
                    multimem_st_u64(mcg + 8 * k, v);
                }
            } else {
                for (int64_t k = s0 + t0; k < s1; k += stride) {
                    unsigned long long v = 0ull;
                    for (int r = 0; r < A.world; ++r)
                        v += __ldcg(reinterpret_cast<const unsigned long long*>(reinterpret_cast<const char*>(A.peers[r]) + kSymmGridOffset) + k);
                    for (int r = 0; r < A.world; ++r)
                        reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(A.peers[r]) + kSymmGridOffset)[k] = v;
                }
            }
        }
        scan_grid_barrier(&A.ws->bar, ++bar_no * G);
        if (b == 0) scan_rank_barrier(A, 2);             // every slice has been broadcast
    }
    scan_grid_barrier(&A.ws->bar, ++bar_no * G);

    // ---- phase D: density and cell centres ---------------------------------------------------------
    scan_finish_phase(D, A.grid, A.density, A.gx, A.gy);
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        if (atomicAdd(&A.ws->exit_ticket, 1u) == (unsigned)G - 1u) {
            A.ws->bar = 0u;
            A.ws->exit_ticket = 0u;
        }
    }
}

// ================================================================================================
// (2) the phases as separate kernels (collectives supplied by the caller between them)
// ================================================================================================
template <class Loader>
__global__ void __launch_bounds__(kScanThreads)
k_scan_bbox(const void* pts, int64_t n, ScanWs* ws, double* packed4) {
    __shared__ double s_red[kScanThreads / 32 * 4];
    __shared__ double s_bb[4];
    __shared__ bool s_last;
    const Loader L{static_cast<decltype(Loader::p)>(pts)};
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    BboxFold<Loader> F;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n; i += 4 * stride) { F.add(L, i); F.add(L, i + stride); F.add(L, i + 2 * stride); F.add(L, i + 3 * stride); }
    for (; i < n; i += stride) F.add(L, i);
    double v[4] = {-F.lo(0), -F.lo(1), F.hi(0), F.hi(1)};
    cta_max4(v, s_red, s_bb);
    double* dpart = reinterpret_cast<double*>(reinterpret_cast<char*>(ws) + sizeof(ScanWs));
    if (threadIdx.x < 4) dpart[(size_t)blockIdx.x * 4 + threadIdx.x] = s_bb[threadIdx.x];
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(&ws->exit_ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    double w[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    for (int q = threadIdx.x; q < (int)gridDim.x; q += blockDim.x) {
#pragma unroll
        for (int c = 0; c < 4; ++c) w[c] = fmax(w[c], __ldcg(dpart + (size_t)q * 4 + c));
    }
    cta_max4(w, s_red, s_bb);
    if (threadIdx.x < 4) packed4[threadIdx.x] = s_bb[threadIdx.x];
    if (threadIdx.x == 0) ws->exit_ticket = 0u;
}

__global__ void k_scan_desc(const double* packed4, double g, int max_nx, int max_ny, int64_t cap_cells, int64_t n_local,
                            lidar_scan_desc* D, int32_t* grid) {
    // one CTA derives the descriptor, the whole grid zeroes the cells
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        double p[4] = {packed4[0], packed4[1], packed4[2], packed4[3]};
        scan_derive_desc(p, g, max_nx, max_ny, cap_cells, n_local, D);
    }
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < cap_cells; k += stride) grid[k] = 0;
}

template <class Loader>
__global__ void __launch_bounds__(kScanThreads)
k_scan_hist(const void* pts, int64_t n, const lidar_scan_desc* Dg, int32_t* grid) {
    __shared__ lidar_scan_desc D;
    if (threadIdx.x == 0) D = *Dg;
    __syncthreads();
    if (D.status != 0) return;
    const Loader L{static_cast<decltype(Loader::p)>(pts)};
    scan_hist_phase(L, n, D, grid);
}

__global__ void __launch_bounds__(kScanThreads)
k_scan_finish(const lidar_scan_desc* Dg, const int32_t* grid, double* density, double* gx, double* gy) {
    __shared__ lidar_scan_desc D;
    if (threadIdx.x == 0) D = *Dg;
    __syncthreads();
    scan_finish_phase(D, grid, density, gx, gy);
}

static size_t scan_ws_bytes() { return ws_align(sizeof(ScanWs) + sizeof(double) * 4 * kScanMaxCtas); }

static int scan_grid_for(int64_t n, int per_sm) {
    int64_t want = (n + kScanThreads * 4 - 1) / (kScanThreads * 4);
    const int64_t cap = (int64_t)sm_count() * per_sm;
    if (want < sm_count()) want = sm_count();
    if (want > cap) want = cap;
    if (want > kScanMaxCtas) want = kScanMaxCtas;
    return (int)want;
}

}  // namespace lidar

using namespace lidar;

extern "C" {

size_t lidar_scan_workspace_bytes(void) { return scan_ws_bytes(); }

size_t lidar_scan_symm_bytes(int64_t cap_cells) {
    if (cap_cells <= 0) return 0;
    return kSymmGridOffset + (size_t)((cap_cells + 3) & ~(int64_t)3) * sizeof(int32_t);
}
size_t lidar_scan_symm_grid_offset(void) { return kSymmGridOffset; }

int lidar_scan_workspace_init(void* d_ws, size_t ws_bytes, void* stream) {
    LIDAR_REQUIRE(d_ws && ws_bytes >= scan_ws_bytes(), LIDAR_ERR_WORKSPACE, "lidar_scan_workspace_init: workspace too small");
    LIDAR_CUDA_TRY(cudaMemsetAsync(d_ws, 0, scan_ws_bytes(), as_stream(stream)));
    return LIDAR_OK;
}

int lidar_scan_density(const void* d_points, int fmt, int64_t n, double grid_size, int max_nx, int max_ny,
                       int64_t cap_cells, int32_t* d_grid, double* d_density, double* d_gx, double* d_gy,
                       lidar_scan_desc* d_desc, lidar_scan_desc* h_desc_mapped, const lidar_scan_comm* comm,
                       uint32_t epoch, void* d_ws, size_t ws_bytes, void* stream) {
    LIDAR_REQUIRE(n >= 0 && (n == 0 || d_points), LIDAR_ERR_INVALID, "lidar_scan_density: bad points");
    LIDAR_REQUIRE(fmt == LIDAR_FMT_F32X4 || fmt == LIDAR_FMT_F64X3, LIDAR_ERR_INVALID, "lidar_scan_density: unknown point format %d", fmt);
    LIDAR_REQUIRE(grid_size > 0.0, LIDAR_ERR_INVALID, "lidar_scan_density: grid_size must be > 0");
    LIDAR_REQUIRE(max_nx > 0 && max_ny > 0 && cap_cells > 0 && cap_cells % 4 == 0 && cap_cells < (1ll << 31), LIDAR_ERR_INVALID,
                  "lidar_scan_density: capacities must be positive (cap_cells a multiple of 4)");
    LIDAR_REQUIRE(d_density && d_gx && d_gy && d_desc, LIDAR_ERR_INVALID, "lidar_scan_density: NULL output");
    LIDAR_REQUIRE(d_ws && ws_bytes >= scan_ws_bytes(), LIDAR_ERR_WORKSPACE, "lidar_scan_density: workspace too small");
    ScanArgs A{};
    A.pts = d_points; A.fmt = fmt; A.n = n; A.g = grid_size; A.max_nx = max_nx; A.max_ny = max_ny; A.cap_cells = cap_cells;
    A.ws = static_cast<ScanWs*>(d_ws);
    A.density = d_density; A.gx = d_gx; A.gy = d_gy; A.desc_out = d_desc; A.desc_host = h_desc_mapped;
    LIDAR_REQUIRE(epoch > 0, LIDAR_ERR_INVALID, "lidar_scan_density: epoch must be > 0 and increase by one per call");
    A.rank = 0; A.world = 1; A.epoch = epoch;
    if (comm && comm->world > 1) {
        LIDAR_REQUIRE(comm->world <= kScanMaxWorld && comm->rank >= 0 && comm->rank < comm->world, LIDAR_ERR_INVALID,
                      "lidar_scan_density: bad communicator (rank %d of %d)", comm->rank, comm->world);
        LIDAR_REQUIRE(comm->symm_bytes >= lidar_scan_symm_bytes(cap_cells), LIDAR_ERR_WORKSPACE,
                      "lidar_scan_density: symmetric buffer too small for cap_cells");
        A.rank = comm->rank; A.world = comm->world;
        for (int r = 0; r < comm->world; ++r) {
            LIDAR_REQUIRE(comm->peer_ptrs[r] != nullptr, LIDAR_ERR_INVALID, "lidar_scan_density: peer %d has no mapping", r);
            A.peers[r] = static_cast<ScanSymmHeader*>(comm->peer_ptrs[r]);
        }
        A.self = A.peers[A.rank];
        A.mc = static_cast<char*>(comm->multicast_ptr);
        A.grid = reinterpret_cast<int32_t*>(reinterpret_cast<char*>(A.self) + kSymmGridOffset);
    } else {
        LIDAR_REQUIRE(d_grid != nullptr, LIDAR_ERR_INVALID, "lidar_scan_density: d_grid is NULL");
        A.grid = d_grid;
    }
    // co-resident persistent grid: as many CTAs per SM as the occupancy calculator allows, capped at 4
    int per_sm = 0;
    const void* fn = fmt == LIDAR_FMT_F32X4 ? (const void*)k_scan_density<LoadF32x4> : (const void*)k_scan_density<LoadF64x3>;
    LIDAR_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, kScanThreads, 0));
    LIDAR_REQUIRE(per_sm >= 1, LIDAR_ERR_CUDA, "lidar_scan_density: kernel does not fit an SM");
    if (per_sm > 4) per_sm = 4;
    const int G = scan_grid_for(n, per_sm);
    void* args[] = {&A};
    LIDAR_CUDA_TRY(cudaLaunchCooperativeKernel(fn, dim3(G), dim3(kScanThreads), args, 0, as_stream(stream)));
    return LIDAR_OK;
}

int lidar_scan_bbox_packed(const void* d_points, int fmt, int64_t n, double* d_packed4, void* d_ws, size_t ws_bytes,
                           void* stream) {
    LIDAR_REQUIRE(n >= 0 && (n == 0 || d_points) && d_packed4, LIDAR_ERR_INVALID, "lidar_scan_bbox_packed: bad argument");
    LIDAR_REQUIRE(d_ws && ws_bytes >= scan_ws_bytes(), LIDAR_ERR_WORKSPACE, "lidar_scan_bbox_packed: workspace too small");
    const int G = scan_grid_for(n, 4);
    if (fmt == LIDAR_FMT_F32X4)
        k_scan_bbox<LoadF32x4><<<G, kScanThreads, 0, as_stream(stream)>>>(d_points, n, static_cast<ScanWs*>(d_ws), d_packed4);
    else if (fmt == LIDAR_FMT_F64X3)
        k_scan_bbox<LoadF64x3><<<G, kScanThreads, 0, as_stream(stream)>>>(d_points, n, static_cast<ScanWs*>(d_ws), d_packed4);
    else
        LIDAR_REQUIRE(false, LIDAR_ERR_INVALID, "lidar_scan_bbox_packed: unknown point format %d", fmt);
    LIDAR_CHECK_LAUNCH();
    return LIDAR_OK;
}

int lidar_scan_hist(const void* d_points, int fmt, int64_t n, const double* d_packed4, double grid_size, int max_nx,
                    int max_ny, int64_t cap_cells, int32_t* d_grid, lidar_scan_desc* d_desc, void* stream) {
    LIDAR_REQUIRE(n >= 0 && (n == 0 || d_points) && d_packed4 && d_grid && d_desc, LIDAR_ERR_INVALID, "lidar_scan_hist: bad argument");
    LIDAR_REQUIRE(grid_size > 0.0 && max_nx > 0 && max_ny > 0 && cap_cells > 0, LIDAR_ERR_INVALID, "lidar_scan_hist: bad sizes");
    cudaStream_t st = as_stream(stream);
    k_scan_desc<<<sm_count(), kScanThreads, 0, st>>>(d_packed4, grid_size, max_nx, max_ny, cap_cells, n, d_desc, d_grid);
    LIDAR_CHECK_LAUNCH();
    if (n > 0) {
        const int G = scan_grid_for(n, 8);
        if (fmt == LIDAR_FMT_F32X4) k_scan_hist<LoadF32x4><<<G, kScanThreads, 0, st>>>(d_points, n, d_desc, d_grid);
        else if (fmt == LIDAR_FMT_F64X3) k_scan_hist<LoadF64x3><<<G, kScanThreads, 0, st>>>(d_points, n, d_desc, d_grid);
        else LIDAR_REQUIRE(false, LIDAR_ERR_INVALID, "lidar_scan_hist: unknown point format %d", fmt);
        LIDAR_CHECK_LAUNCH();
    }
    return LIDAR_OK;
}

int lidar_scan_finish(const int32_t* d_grid, const lidar_scan_desc* d_desc, double* d_density, double* d_gx, double* d_gy,
                      void* stream) {
    LIDAR_REQUIRE(d_grid && d_desc && d_density && d_gx && d_gy, LIDAR_ERR_INVALID, "lidar_scan_finish: NULL argument");
    k_scan_finish<<<sm_count() * 4, kScanThreads, 0, as_stream(stream)>>>(d_desc, d_grid, d_density, d_gx, d_gy);
    LIDAR_CHECK_LAUNCH();
    return LIDAR_OK;
}

}  // extern "C"
