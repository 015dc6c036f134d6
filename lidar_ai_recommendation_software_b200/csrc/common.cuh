// Shared device/host helpers for the lidar_b200 CUDA core (sm_100a only).
//
// Conventions used by every kernel in this directory:
//  * point clouds arrive in one of two layouts (LIDAR_FMT_* in include/lidar_b200.h):
//      F32X4  float4 (x, y, z, intensity), 16-byte aligned  -> one LDG.128 per point
//      F64X3  packed (n,3) float64 rows, the reference's own layout
//             (utils/data_processing.py:34-41 keeps only xyz as float64)
//  * every integer-valued decision (bin index, voxel index, mask bit, neighbour test) is
//    taken in fp64 on the exact widening of the stored value, with explicit
//    __dadd_rn/__dmul_rn/__ddiv_rn so that ptxas cannot contract a*b+c into an FMA
//    (SURVEY.md Appendix A.2/A.3).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/lidar_b200.h"

namespace lidar {

// ---------------------------------------------------------------------------------------------
// error plumbing (no exceptions cross the C ABI)
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define LIDAR_CUDA_TRY(expr)                                                      \
    do {                                                                          \
        cudaError_t _e = (expr);                                                  \
        if (_e != cudaSuccess) return ::lidar::cuda_fail(_e, #expr, __FILE__, __LINE__); \
    } while (0)

#define LIDAR_CHECK_LAUNCH() LIDAR_CUDA_TRY(cudaGetLastError())

#define LIDAR_REQUIRE(cond, code, ...)        \
    do {                                      \
        if (!(cond)) {                        \
            ::lidar::set_error(__VA_ARGS__);  \
            return (code);                    \
        }                                     \
    } while (0)

int sm_count();          // SMs of the current device (148 on B200), cached per device
size_t smem_optin();     // max opt-in dynamic shared memory per block

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// reduce.cu: lidar_moments of an (n,3) fp64 cloud whose row count / centre live in device memory (either may be
// NULL: then cap rows / centre 0).  The workspace's ticket must be zero (it is after every reduction).
int moments_f64x3_dev(const double* d_points, int64_t cap, const long long* d_n, const double* d_center3,
                      double* d_out6, void* d_reduce_ws, cudaStream_t st);

// Bump allocator over the caller-provided workspace (no hidden cudaMalloc in the hot path).
struct Workspace {
    char* base;
    size_t size;
    size_t off;
    Workspace(void* p, size_t n) : base(static_cast<char*>(p)), size(n), off(0) {}
    template <class T>
    T* take(size_t count) {
        size_t a = (off + 255) & ~size_t(255);
        size_t bytes = count * sizeof(T);
        if (base == nullptr || a + bytes > size) {
            off = size + 1;  // poison
            return nullptr;
        }
        off = a + bytes;
        return reinterpret_cast<T*>(base + a);
    }
    bool ok() const { return off <= size; }
};
static inline size_t ws_align(size_t b) { return (b + 255) & ~size_t(255); }

// ---------------------------------------------------------------------------------------------
// point loaders
// ---------------------------------------------------------------------------------------------
struct Pt {
    double x, y, z, w;
};

struct LoadF32x4 {
    const float4* __restrict__ p;
    static constexpr bool kHasW = true;
    __device__ __forceinline__ float4 raw(int64_t i) const {
        // streaming read: every point is touched once per pass, do not pollute L1
        float4 v;
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                     : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                     : "l"(p + i));
        return v;
    }
    __device__ __forceinline__ Pt load(int64_t i) const {
        float4 v = raw(i);
        return Pt{(double)v.x, (double)v.y, (double)v.z, (double)v.w};
    }
};

struct LoadF64x3 {
    const double* __restrict__ p;
    static constexpr bool kHasW = false;
    __device__ __forceinline__ Pt load(int64_t i) const {
        const double* q = p + 3 * i;
        return Pt{__ldg(q), __ldg(q + 1), __ldg(q + 2), 0.0};
    }
};

// two independent strided fp64 columns (e.g. points[:, :2] of an (n,3) array, or an (m,2) array)
struct LoadF64uv {
    const double* __restrict__ u;
    const double* __restrict__ v;
    int64_t su, sv;
    static constexpr bool kHasW = false;
    __device__ __forceinline__ Pt load(int64_t i) const {
        return Pt{__ldg(u + i * su), __ldg(v + i * sv), 0.0, 0.0};
    }
};

// ---------------------------------------------------------------------------------------------
// warp helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ unsigned warp_sum_u32(unsigned v) {
    return __reduce_add_sync(0xffffffffu, v);
}

// ---------------------------------------------------------------------------------------------
// single-pass chained scan ("decoupled look-back") over tile aggregates.
//   desc[t]: bits 63..62 = state (0 empty, 1 aggregate only, 2 inclusive prefix), bits 61..0 value.
// Called by ALL 32 lanes of one warp of the CTA that owns logical tile `tile`; tiles are handed
// out through an atomic ticket so that tile t-1 is always already running (forward progress).
// Returns the exclusive prefix of the tile in every lane.
// ---------------------------------------------------------------------------------------------
constexpr unsigned long long kScanAgg = 1ull << 62;
constexpr unsigned long long kScanInc = 2ull << 62;
constexpr unsigned long long kScanValMask = (1ull << 62) - 1;

__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ unsigned long long scan_lookback_warp(unsigned long long* desc, int tile,
                                                                 unsigned long long aggregate) {
    const unsigned lane = lane_id();
    if (tile == 0) {
        if (lane == 0) st_relaxed_u64(desc, kScanInc | aggregate);
        return 0ull;
    }
    if (lane == 0) st_relaxed_u64(desc + tile, kScanAgg | aggregate);
    unsigned long long exclusive = 0ull;
    int look = tile - 1;  // lane L inspects tile look - L
    while (true) {
        const int t = look - (int)lane;
        unsigned long long d = kScanInc;  // virtual tiles < 0: inclusive prefix 0
        if (t >= 0) {
            do {
                d = ld_relaxed_u64(desc + t);
            } while ((d >> 62) == 0ull);
        }
        const unsigned inc_mask = __ballot_sync(0xffffffffu, (d >> 62) == 2ull);
        // lanes closer than (and including) the first inclusive-prefix lane contribute
        unsigned long long contrib = d & kScanValMask;
        const int first_inc = inc_mask ? (__ffs(inc_mask) - 1) : 32;
        if ((int)lane > first_inc) contrib = 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
        exclusive += contrib;
        if (inc_mask) break;
        look -= 32;
    }
    if (lane == 0) st_relaxed_u64(desc + tile, kScanInc | ((exclusive + aggregate) & kScanValMask));
    return exclusive;
}

// fp64 helpers that must never be contracted
__device__ __forceinline__ double edge_at(const double* __restrict__ e, int i) { return e[i]; }

}  // namespace lidar
