// K2-K4 — the per-point stages of the reference's preprocess functions on (n,3) float64 clouds:
//
//   lidar_sigma_filter     3-sigma inlier mask + order-preserving compaction + height colours
//                          utils/data_processing.py:143-157 ; app_simplified.py:80-91
//   lidar_select_kth       two adjacent order statistics of a strided fp64 column (radix select),
//                          the inputs of np.percentile(z, 30)   data_processing.py:164 ; app_simplified.py:98
//   lidar_ground_split     z <= thr split: plane-fit moments over the ground points and compaction of
//                          the non-ground points (+ their inlier indices)   data_processing.py:165-188
//   lidar_standardize      (x - mean) / scale, StandardScaler.transform     data_processing.py:190-191
//   lidar_scatter_labels   full_labels = -1; full_labels[non_ground] = labels   data_processing.py:203-204
//
// All of them stream the cloud once (24 B/point in, <= 48 B/point out): HBM-bound.  Every compare
// that decides a mask bit is done in fp64 on the stored values with the reference's own expression.
#include "compact.cuh"

namespace lidar {

// order-preserving map fp64 -> u64 (radix select keys; bbox atomics of the chained preprocess)
__device__ __forceinline__ unsigned long long f64_key(double d) {
    unsigned long long b = (unsigned long long)__double_as_longlong(d);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_f64(unsigned long long k) {
    unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}

// ------------------------------------------------------------------------------------------------
// 3-sigma filter
// ------------------------------------------------------------------------------------------------
struct SigmaParams {
    double mean[3], thr[3], tol[3];
    double zmin, zden;  // colours: h = (z - zmin) / zden
};

__global__ void __launch_bounds__(kCmpThreads)
sigma_filter_kernel(const double* __restrict__ pts, int64_t n, SigmaParams P, uint8_t* __restrict__ mask,
                    double* __restrict__ out_pts, double* __restrict__ out_col, int64_t* __restrict__ count,
                    unsigned long long* __restrict__ guard, unsigned long long* tile_desc, CompactCtrl* ctrl,
                    int n_tiles, const lidar_front_desc* __restrict__ F = nullptr) {
    __shared__ int s_tile;
    if (F) {                                  // chained preprocess: the statistics were left on the device
#pragma unroll
        for (int c = 0; c < 3; ++c) { P.mean[c] = F->mean[c]; P.thr[c] = F->thr[c]; P.tol[c] = F->tol[c]; }
        P.zmin = F->zmin; P.zden = F->zden;
    }
    unsigned guard_local = 0;
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
                const double ax = fabs(__dsub_rn(x[j], P.mean[0]));
                const double ay = fabs(__dsub_rn(y[j], P.mean[1]));
                const double az = fabs(__dsub_rn(z[j], P.mean[2]));
                keep[j] = ax < P.thr[0] && ay < P.thr[1] && az < P.thr[2];
                // knife-edge certificate: mean/std come from a parallel fp64 reduction and may differ
                // from numpy's sequential sum in the last bits; the mask can only differ for points
                // this close to the threshold
                if (fabs(ax - P.thr[0]) <= P.tol[0] || fabs(ay - P.thr[1]) <= P.tol[1] ||
                    fabs(az - P.thr[2]) <= P.tol[2])
                    ++guard_local;
                if (mask) mask[i] = keep[j] ? 1 : 0;
            }
        }
        long long slot[kCmpRows];
        compact_slots(keep, tile, tile_desc, slot, reinterpret_cast<unsigned long long*>(count), tile == n_tiles - 1);
#pragma unroll
        for (int j = 0; j < kCmpRows; ++j)
            if (slot[j] >= 0) {
                double* o = out_pts + 3 * slot[j];
                o[0] = x[j]; o[1] = y[j]; o[2] = z[j];
                if (out_col) {
                    const double h = __ddiv_rn(__dsub_rn(z[j], P.zmin), P.zden);
                    double* c = out_col + 3 * slot[j];
                    c[0] = h;
                    c[1] = __dmul_rn(0.5, __dsub_rn(1.0, h));
                    c[2] = 0.5;
                }
            }
    }
    if (guard_local) atomicAdd(guard, (unsigned long long)guard_local);
}

// ------------------------------------------------------------------------------------------------
// ground split: plane moments over z <= thr, compaction of z > thr
// ------------------------------------------------------------------------------------------------
constexpr int kPlaneMaxBlocks = 1024;
struct PlaneWs {
    double partial[kPlaneMaxBlocks][10];
    unsigned int ticket;
};

__global__ void __launch_bounds__(kCmpThreads, 2)
ground_split_kernel(const double* __restrict__ pts, int64_t n, double thr, double cx, double cy, double cz,
                    double* __restrict__ out_pts, int32_t* __restrict__ out_index, int64_t* __restrict__ count,
                    double* __restrict__ plane_out10, unsigned long long* __restrict__ guard, double tol,
                    unsigned long long* tile_desc, CompactCtrl* ctrl, PlaneWs* pw, int n_tiles,
                    lidar_front_desc* F = nullptr) {
    __shared__ int s_tile;
    if (F) {                                  // chained preprocess: count, threshold and centre from the device
        n = F->n_in; thr = F->z_thr; cx = F->mean[0]; cy = F->mean[1]; cz = F->mean[2];
    }
    // {n, Sx, Sy, Sz, Sxx, Sxy, Syy, Sxz, Syz, unused} about the centre (cx,cy,cz)
    double s[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    // chained preprocess only: bbox of all rows [0..2] min, [3..5] max, and of the rows kept [6..11]
    double bb[12] = {INFINITY, INFINITY, INFINITY, -INFINITY, -INFINITY, -INFINITY,
                     INFINITY, INFINITY, INFINITY, -INFINITY, -INFINITY, -INFINITY};
    unsigned guard_local = 0;
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
                const bool ground = z[j] <= thr;
                keep[j] = !ground;
                if (fabs(z[j] - thr) <= tol && z[j] != thr) ++guard_local;
                if (F) {
                    bb[0] = fmin(bb[0], x[j]); bb[1] = fmin(bb[1], y[j]); bb[2] = fmin(bb[2], z[j]);
                    bb[3] = fmax(bb[3], x[j]); bb[4] = fmax(bb[4], y[j]); bb[5] = fmax(bb[5], z[j]);
                    if (!ground) {
                        bb[6] = fmin(bb[6], x[j]); bb[7] = fmin(bb[7], y[j]); bb[8] = fmin(bb[8], z[j]);
                        bb[9] = fmax(bb[9], x[j]); bb[10] = fmax(bb[10], y[j]); bb[11] = fmax(bb[11], z[j]);
                    }
                }
                if (ground) {
                    const double dx = x[j] - cx, dy = y[j] - cy, dz = z[j] - cz;
                    s[0] += 1.0; s[1] += dx; s[2] += dy; s[3] += dz;
                    s[4] += dx * dx; s[5] += dx * dy; s[6] += dy * dy; s[7] += dx * dz; s[8] += dy * dz;
                }
            }
        }
        long long slot[kCmpRows];
        compact_slots(keep, tile, tile_desc, slot, reinterpret_cast<unsigned long long*>(count), tile == n_tiles - 1);
#pragma unroll
        for (int j = 0; j < kCmpRows; ++j)
            if (slot[j] >= 0) {
                double* o = out_pts + 3 * slot[j];
                o[0] = x[j]; o[1] = y[j]; o[2] = z[j];
                out_index[slot[j]] = (int32_t)(base + (int64_t)j * kCmpThreads + threadIdx.x);
            }
    }
    if (guard_local) atomicAdd(guard, (unsigned long long)guard_local);
    if (F) {                                  // min / max select, so ordered-key atomics are exact and order free
        __shared__ double s_bb[kCmpThreads / 32][12];
#pragma unroll
        for (int c = 0; c < 12; ++c) {
            const bool is_max = (c % 6) >= 3;
            bb[c] = is_max ? warp_max(bb[c]) : warp_min(bb[c]);
        }
        if (lane_id() == 0) {
#pragma unroll
            for (int c = 0; c < 12; ++c) s_bb[threadIdx.x >> 5][c] = bb[c];
        }
        __syncthreads();
        if (threadIdx.x < 12) {               // one atomic per CTA and channel
            const int c = threadIdx.x;
            const bool is_max = (c % 6) >= 3;
            double v = s_bb[0][c];
            for (int w = 1; w < kCmpThreads / 32; ++w) v = is_max ? fmax(v, s_bb[w][c]) : fmin(v, s_bb[w][c]);
            if (v != (is_max ? -INFINITY : INFINITY)) {
                unsigned long long* dst = reinterpret_cast<unsigned long long*>(c < 6 ? F->key_in : F->key_ng) + (c % 6);
                if (is_max) atomicMax(dst, f64_key(v)); else atomicMin(dst, f64_key(v));
            }
        }
    }
    // deterministic fold of the plane moments: warp tree, CTA partial, last CTA folds in CTA order
    __shared__ double s_p[kCmpThreads / 32][9];
    __shared__ bool s_last;
    const int warp = threadIdx.x >> 5;
#pragma unroll
    for (int c = 0; c < 9; ++c) s[c] = warp_sum(s[c]);
    if (lane_id() == 0) {
#pragma unroll
        for (int c = 0; c < 9; ++c) s_p[warp][c] = s[c];
    }
    __syncthreads();
    if (threadIdx.x < 9) {
        double v = 0;
        for (int w = 0; w < kCmpThreads / 32; ++w) v += s_p[w][threadIdx.x];
        pw->partial[blockIdx.x][threadIdx.x] = v;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&pw->ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (threadIdx.x < 9) {
        double v = 0;
        for (unsigned b = 0; b < gridDim.x; ++b) v += __ldcg(&pw->partial[b][threadIdx.x]);
        plane_out10[threadIdx.x] = v;
    }
    if (threadIdx.x == 0) { plane_out10[9] = 0.0; pw->ticket = 0u; }
    if (F && threadIdx.x < 12) {              // every CTA's atomics are visible: its ticket came after a fence
        const unsigned long long* src = reinterpret_cast<const unsigned long long*>(threadIdx.x < 6 ? F->key_in : F->key_ng) + (threadIdx.x % 6);
        (threadIdx.x < 6 ? F->bbox_in : F->bbox_ng)[threadIdx.x % 6] = key_f64(*((volatile const unsigned long long*)src));
    }
}

// ------------------------------------------------------------------------------------------------
// radix select on fp64 keys (8 passes of 8 bits), state kept on the device.  Passes 0 and 1 read the column; the keys
// that carry the winning 16-bit prefix (a few per cent of a height column) are then compacted, and passes 2..7 run on
// that buffer instead of re-reading the whole column six more times.
// ------------------------------------------------------------------------------------------------
struct SelectState {
    unsigned long long prefix;   // key bits fixed so far (high bits)
    long long k;                 // rank still to find inside the current bucket
    unsigned int hist[256];
    unsigned int ticket;
    unsigned int pad;
    // second order statistic
    unsigned long long count_le; // #elements <= kth
    unsigned long long min_gt;   // smallest key > kth (key space)
    unsigned long long n_buf;    // keys that survived the first two passes (compacted, see select_compact_kernel)
    long long k0;                // the rank asked for (st->k is consumed by the passes)
};


constexpr int kSelThreads = 256;

// Last CTA of a pass: the digit whose bucket holds rank k.  One thread per digit and a block-wide scan of the 256
// counts (a single thread walking them with dependent loads took ~10 us per pass -- most of a pass on 1 M keys).
__device__ __forceinline__ void select_pick_digit(SelectState* st, unsigned long long prefix) {
    __shared__ unsigned s_w[kSelThreads / 32];
    const unsigned c = ((volatile unsigned*)st->hist)[threadIdx.x];     // kSelThreads == 256 == digits
    const long long k = st->k;
    unsigned inc = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane_id() >= (unsigned)o) inc += t;
    }
    if (lane_id() == 31) s_w[threadIdx.x >> 5] = inc;
    __syncthreads();                                  // also: every thread has read st->k before anyone writes it
    unsigned off = 0;
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) off += s_w[w];
    const long long hi = (long long)off + inc, lo = hi - c;
    if ((k >= lo && k < hi) || (threadIdx.x == 255 && k >= hi)) {
        st->k = k >= hi ? k - hi : k - lo;
        st->prefix = (prefix << 8) | (unsigned long long)threadIdx.x;
    }
    st->hist[threadIdx.x] = 0;
    if (threadIdx.x == 0) st->ticket = 0;
}

__global__ void __launch_bounds__(kSelThreads)
select_pass_kernel(const double* __restrict__ col, int64_t stride_el, int64_t n, int pass, SelectState* st,
                   const long long* __restrict__ d_n = nullptr) {
    if (d_n) n = *d_n;
    __shared__ unsigned s_hist[256];
    __shared__ bool s_last;
    s_hist[threadIdx.x] = 0;  // kSelThreads == 256
    __syncthreads();
    const unsigned long long prefix = st->prefix;
    const int shift = 56 - 8 * pass;
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
        const unsigned long long key = f64_key(__ldg(col + i * stride_el));
        const bool match = pass == 0 ? true : ((key >> (shift + 8)) == prefix);
        if (match) atomicAdd(&s_hist[(key >> shift) & 0xff], 1u);
    }
    __syncthreads();
    if (s_hist[threadIdx.x]) atomicAdd(&st->hist[threadIdx.x], s_hist[threadIdx.x]);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&st->ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    select_pick_digit(st, prefix);
}

// keys whose top 16 bits equal the prefix fixed by passes 0 and 1 -> buf (order irrelevant: later passes only count)
__global__ void __launch_bounds__(kSelThreads)
select_compact_kernel(const double* __restrict__ col, int64_t stride_el, int64_t n, SelectState* st,
                      unsigned long long* __restrict__ buf, const long long* __restrict__ d_n = nullptr) {
    if (d_n) n = *d_n;
    const unsigned long long prefix = st->prefix;     // 16 bits
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    const int64_t n_round = ((n + 31) / 32) * 32;     // whole warps stay in the loop for the ballot
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += step) {
        unsigned long long key = 0ull;
        bool keep = false;
        if (i < n) {
            key = f64_key(__ldg(col + i * stride_el));
            keep = (key >> 48) == prefix;
        }
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (m) {
            unsigned long long base = 0ull;
            if (lane_id() == 0) base = atomicAdd(&st->n_buf, (unsigned long long)__popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (keep) buf[base + __popc(m & lanemask_lt())] = key;
        }
    }
}

// passes 2..7 on the compacted keys
__global__ void __launch_bounds__(kSelThreads)
select_pass_buf_kernel(const unsigned long long* __restrict__ buf, int pass, SelectState* st) {
    __shared__ unsigned s_hist[256];
    __shared__ bool s_last;
    s_hist[threadIdx.x] = 0;  // kSelThreads == 256
    __syncthreads();
    const unsigned long long prefix = st->prefix;
    const int64_t n = (int64_t)st->n_buf;
    const int shift = 56 - 8 * pass;
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
        const unsigned long long key = __ldcg(buf + i);
        if ((key >> (shift + 8)) == prefix) atomicAdd(&s_hist[(key >> shift) & 0xff], 1u);
    }
    __syncthreads();
    if (s_hist[threadIdx.x]) atomicAdd(&st->hist[threadIdx.x], s_hist[threadIdx.x]);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&st->ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    select_pick_digit(st, prefix);
}

__global__ void __launch_bounds__(kSelThreads)
select_next_kernel(const double* __restrict__ col, int64_t stride_el, int64_t n, SelectState* st,
                   double* __restrict__ out2, lidar_front_desc* F = nullptr) {
    if (F) n = F->n_in;
    __shared__ unsigned long long s_cnt[kSelThreads / 32];
    __shared__ unsigned long long s_min[kSelThreads / 32];
    __shared__ bool s_last;
    const unsigned long long kth = st->prefix;
    unsigned long long cnt = 0, mn = ~0ull;
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
        const unsigned long long key = f64_key(__ldg(col + i * stride_el));
        if (key <= kth) ++cnt;
        else if (key < mn) mn = key;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        const unsigned long long t = __shfl_xor_sync(0xffffffffu, mn, o);
        mn = t < mn ? t : mn;
    }
    if (lane_id() == 0) { s_cnt[threadIdx.x >> 5] = cnt; s_min[threadIdx.x >> 5] = mn; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kSelThreads / 32; ++w) { cnt += s_cnt[w]; mn = s_min[w] < mn ? s_min[w] : mn; }
        atomicAdd(&st->count_le, cnt);
        atomicMin(&st->min_gt, mn);
        __threadfence();
        s_last = (atomicAdd(&st->ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last || threadIdx.x != 0) return;
    __threadfence();
    const unsigned long long cle = *((volatile unsigned long long*)&st->count_le);
    const unsigned long long mgt = *((volatile unsigned long long*)&st->min_gt);
    const long long k = st->k0;
    const double a = key_f64(kth);
    const double b = ((long long)cle >= k + 2 || mgt == ~0ull) ? key_f64(kth) : key_f64(mgt);
    out2[0] = a;
    out2[1] = b;
    st->ticket = 0;
    if (F) {
        // np.percentile(z, 30), method 'linear' (numpy/lib/_function_base_impl.py _quantile + _lerp): virtual index
        // (n-1)*q, t = its fraction, a + (b-a)*t below one half and b - (b-a)*(1-t) from there on
        const double virt = __dmul_rn((double)(n - 1), __ddiv_rn(30.0, 100.0));
        const double t = __dsub_rn(virt, floor(virt));
        const double d = __dsub_rn(b, a);
        F->z_thr = t >= 0.5 ? __dsub_rn(b, __dmul_rn(d, __dsub_rn(1.0, t))) : __dadd_rn(a, __dmul_rn(d, t));
    }
}

__global__ void select_init_kernel(SelectState* st, int64_t k, const lidar_front_desc* F = nullptr) {
    if (threadIdx.x < 256) st->hist[threadIdx.x] = 0;
    if (threadIdx.x == 0) {
        if (F) {                              // lo = floor((n_in - 1) * 0.3), the lower rank of the 30th percentile
            const long long n_in = F->n_in;
            k = n_in > 0 ? (long long)floor(__dmul_rn((double)(n_in - 1), __ddiv_rn(30.0, 100.0))) : 0;
        }
        st->k0 = k;
        st->prefix = 0; st->k = k; st->ticket = 0; st->pad = 0; st->count_le = 0; st->min_gt = ~0ull; st->n_buf = 0ull;
    }
}

// ------------------------------------------------------------------------------------------------
// elementwise helpers
// ------------------------------------------------------------------------------------------------
__global__ void standardize_kernel(const double* __restrict__ in, int64_t n3, double m0, double m1, double m2,
                                   double s0, double s1, double s2, double* __restrict__ out) {
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n3; i += step) {
        const int c = (int)(i % 3);
        const double m = c == 0 ? m0 : (c == 1 ? m1 : m2);
        const double s = c == 0 ? s0 : (c == 1 ? s1 : s2);
        out[i] = __ddiv_rn(__dsub_rn(in[i], m), s);   // X -= mean; X /= scale
    }
}

__global__ void scatter_labels_kernel(const int32_t* __restrict__ labels, const int32_t* __restrict__ index,
                                      int64_t m, long long* __restrict__ full) {
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += step)
        full[index[i]] = (long long)labels[i];
}
__global__ void fill_i64_kernel(long long* __restrict__ p, int64_t n, long long v) {
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) p[i] = v;
}

// out[i] = in[idx[i]] for rows of `row_words` 4-byte words (downsample_point_cloud's fancy index)
__global__ void gather_rows_kernel(const uint32_t* __restrict__ in, int row_words, const long long* __restrict__ idx,
                                   int64_t k, uint32_t* __restrict__ out) {
    const int64_t total = k * row_words;
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += step) {
        const int64_t r = t / row_words;
        const int w = (int)(t - r * row_words);
        out[t] = in[idx[r] * row_words + w];
    }
}

// ------------------------------------------------------------------------------------------------
// chained preprocess (lidar_preprocess_front): the scalar glue between the stages, one thread each.  The
// expressions are the ones preprocess.py evaluates with numpy on the host path (same IEEE operations, same order).
// ------------------------------------------------------------------------------------------------
enum { kGlueInit = 0, kGlueMean, kGlueStd, kGlueScMean, kGlueScale, kGlueXMean, kGlueEps };

__global__ void front_glue_kernel(lidar_front_desc* F, int stage, long long n) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double m = (double)F->n_nonground;
    switch (stage) {
    case kGlueInit:
        for (int c = 0; c < 3; ++c) { F->key_in[c] = F->key_ng[c] = ~0ull; F->key_in[3 + c] = F->key_ng[3 + c] = 0ull; }
        break;
    case kGlueMean:                                        // np.mean(points, axis=0): sum / n
        for (int c = 0; c < 3; ++c) F->mean[c] = __ddiv_rn(F->sum1[c], (double)n);
        break;
    case kGlueStd:                                         // np.std: sqrt(mean of squared deviations); 3 sigma
        for (int c = 0; c < 3; ++c) {
            const double sd = __dsqrt_rn(__ddiv_rn(F->sum2[3 + c], (double)n));
            F->std[c] = sd;
            F->thr[c] = __dmul_rn(3.0, sd);
            F->tol[c] = __dmul_rn(1e-9, sd);
        }
        F->zmin = F->bbox_raw[2];
        F->zden = __dadd_rn(__dsub_rn(F->bbox_raw[6], F->bbox_raw[2]), 1e-10);
        break;
    case kGlueScMean:                                      // StandardScaler.fit: mean_
        for (int c = 0; c < 3; ++c) F->sc_mean[c] = __ddiv_rn(F->t1[c], m);
        break;
    case kGlueScale:                                       // sklearn _incremental_mean_and_var + _handle_zeros_in_scale
        for (int c = 0; c < 3; ++c) {
            const double corr = __ddiv_rn(__dmul_rn(F->t2[c], F->t2[c]), m);
            double var = __ddiv_rn(__dsub_rn(F->t2[3 + c], corr), m);
            if (!(var > 0.0)) var = 0.0;                   // rounding can leave a tiny negative variance: sqrt would be NaN
            // StandardScaler.fit: constant_mask = _is_constant_feature(var_, mean_, n_samples_seen_), i.e.
            //   var <= n*eps*var + (n*mean*eps)^2  (a near-constant column with a large mean, e.g. UTM x of one scan
            // line, counts as constant), then scale_ = _handle_zeros_in_scale(sqrt(var_), constant_mask=constant_mask)
            const double eps = 2.220446049250313e-16;
            const double nme = __dmul_rn(__dmul_rn(m, F->sc_mean[c]), eps);
            const double bound = __dadd_rn(__dmul_rn(__dmul_rn(m, eps), var), __dmul_rn(nme, nme));
            double sc = __dsqrt_rn(var);
            if (var <= bound) sc = 1.0;
            F->scale[c] = sc;
        }
        break;
    case kGlueXMean:
        for (int c = 0; c < 3; ++c) F->xm[c] = __ddiv_rn(F->u1[c], m);
        break;
    case kGlueEps: {                                       // eps = max(0.2, min(0.5, np.mean(np.std(X, axis=0)) * 0.5))
        for (int c = 0; c < 3; ++c) F->xstd[c] = __dsqrt_rn(__ddiv_rn(F->u2[3 + c], m));
        const double avg = __dmul_rn(__ddiv_rn(__dadd_rn(__dadd_rn(F->xstd[0], F->xstd[1]), F->xstd[2]), 3.0), 0.5);
        double e = avg < 0.5 ? avg : 0.5;
        e = e > 0.2 ? e : 0.2;
        F->eps = e;
        break;
    }
    default: break;
    }
}

__global__ void standardize_front_kernel(const double* __restrict__ in, const lidar_front_desc* __restrict__ F,
                                         double* __restrict__ out) {
    const int64_t n3 = 3 * F->n_nonground;
    const double m0 = F->sc_mean[0], m1 = F->sc_mean[1], m2 = F->sc_mean[2];
    const double s0 = F->scale[0], s1 = F->scale[1], s2 = F->scale[2];
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n3; i += step) {
        const int c = (int)(i % 3);
        const double m = c == 0 ? m0 : (c == 1 ? m1 : m2);
        const double s = c == 0 ? s0 : (c == 1 ? s1 : s2);
        out[i] = __ddiv_rn(__dsub_rn(in[i], m), s);
    }
}

static int ew_grid(int64_t n) {
    int64_t want = (n + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (want < 1) want = 1;
    return (int)(want < cap ? want : cap);
}

struct PreLayout {
    size_t off_ctrl, off_desc, off_plane, off_guard, total;
    int64_t tiles;
};
static PreLayout pre_layout(int64_t n) {
    PreLayout L;
    const CompactLayout C = compact_layout(n);
    L.tiles = C.tiles;
    L.off_ctrl = C.off_ctrl;
    L.off_desc = C.off_desc;
    L.off_plane = ws_align(C.total);
    L.off_guard = ws_align(L.off_plane + sizeof(PlaneWs));
    L.total = ws_align(L.off_guard + 256);
    return L;
}

struct FrontLayout {
    size_t off_reduce, off_pre, off_select, total;
};
static FrontLayout front_layout(int64_t n) {
    FrontLayout L;
    L.off_reduce = 0;
    L.off_pre = ws_align(lidar_reduce_workspace_bytes());
    L.off_select = ws_align(L.off_pre + pre_layout(n).total);
    L.total = ws_align(L.off_select + ws_align(sizeof(SelectState)) + sizeof(unsigned long long) * (size_t)n);
    return L;
}

}  // namespace lidar

using namespace lidar;

extern "C" {

size_t lidar_preprocess_front_workspace_bytes(int64_t n) { return front_layout(n < 0 ? 0 : n).total; }

int lidar_preprocess_front(const double* d_points, int64_t n, int flags, double* d_inliers, double* d_colors,
                           double* d_nonground, int32_t* d_ng_index, double* d_scaled, lidar_front_desc* d_front,
                           void* d_ws, size_t ws_bytes, void* stream) {
    LIDAR_REQUIRE(n > 0 && n < (1ll << 31), LIDAR_ERR_INVALID, "lidar_preprocess_front: need 0 < n < 2^31");
    LIDAR_REQUIRE(d_points && d_inliers && d_nonground && d_ng_index && d_front, LIDAR_ERR_INVALID,
                  "lidar_preprocess_front: NULL argument");
    LIDAR_REQUIRE(!(flags & LIDAR_FRONT_COLORS) || d_colors, LIDAR_ERR_INVALID, "lidar_preprocess_front: colours without a buffer");
    LIDAR_REQUIRE(!(flags & LIDAR_FRONT_SCALER) || d_scaled, LIDAR_ERR_INVALID, "lidar_preprocess_front: scaler without a buffer");
    const FrontLayout FL = front_layout(n);
    LIDAR_REQUIRE(d_ws && ws_bytes >= FL.total, LIDAR_ERR_WORKSPACE, "lidar_preprocess_front: workspace too small (%zu < %zu)",
                  ws_bytes, FL.total);
    cudaStream_t st = as_stream(stream);
    char* ws = static_cast<char*>(d_ws);
    void* rws = ws + FL.off_reduce;
    char* pws = ws + FL.off_pre;
    const PreLayout L = pre_layout(n);
    lidar_front_desc* F = d_front;
    auto glue = [&](int stage) -> int {
        front_glue_kernel<<<1, 32, 0, st>>>(F, stage, (long long)n);
        LIDAR_CHECK_LAUNCH();
        return LIDAR_OK;
    };
    auto reset_compaction = [&]() -> int {
        LIDAR_CUDA_TRY(cudaMemsetAsync(pws, 0, ws_align(L.off_desc + sizeof(unsigned long long) * L.tiles), st));
        return LIDAR_OK;
    };
    int rc;
    LIDAR_CUDA_TRY(cudaMemsetAsync(F, 0, sizeof(lidar_front_desc), st));
    LIDAR_CUDA_TRY(cudaMemsetAsync(rws, 0, lidar_reduce_workspace_bytes(), st));
    if ((rc = glue(kGlueInit)) != LIDAR_OK) return rc;
    // raw cloud: bbox, mean, population std
    if ((rc = lidar_bbox(d_points, LIDAR_FMT_F64X3, n, F->bbox_raw, rws, lidar_reduce_workspace_bytes(), stream)) != LIDAR_OK) return rc;
    if ((rc = moments_f64x3_dev(d_points, n, nullptr, nullptr, F->sum1, rws, st)) != LIDAR_OK) return rc;
    if ((rc = glue(kGlueMean)) != LIDAR_OK) return rc;
    if ((rc = moments_f64x3_dev(d_points, n, nullptr, F->mean, F->sum2, rws, st)) != LIDAR_OK) return rc;
    if ((rc = glue(kGlueStd)) != LIDAR_OK) return rc;
    // 3-sigma filter
    int grid = sm_count() * 4;
    if ((int64_t)grid > L.tiles) grid = (int)L.tiles;
    if ((rc = reset_compaction()) != LIDAR_OK) return rc;
    sigma_filter_kernel<<<grid, kCmpThreads, 0, st>>>(d_points, n, SigmaParams{}, nullptr, d_inliers,
        (flags & LIDAR_FRONT_COLORS) ? d_colors : nullptr, &F->n_in, reinterpret_cast<unsigned long long*>(&F->guard_sigma),
        reinterpret_cast<unsigned long long*>(pws + L.off_desc), reinterpret_cast<CompactCtrl*>(pws + L.off_ctrl),
        (int)L.tiles, F);
    LIDAR_CHECK_LAUNCH();
    // 30th percentile of the inlier heights: radix select with the count and the rank read on the device
    {
        SelectState* S = reinterpret_cast<SelectState*>(ws + FL.off_select);
        unsigned long long* buf = reinterpret_cast<unsigned long long*>(ws + FL.off_select + ws_align(sizeof(SelectState)));
        const double* col = d_inliers + 2;
        const long long* d_n = reinterpret_cast<const long long*>(&F->n_in);
        select_init_kernel<<<1, 256, 0, st>>>(S, 0, F);
        LIDAR_CHECK_LAUNCH();
        int sgrid = (int)((n + kSelThreads * 8 - 1) / (kSelThreads * 8));
        const int cap = sm_count() * 8;
        sgrid = sgrid < 1 ? 1 : (sgrid > cap ? cap : sgrid);
        if (n >= 65536) {
            for (int pass = 0; pass < 2; ++pass) {
                select_pass_kernel<<<sgrid, kSelThreads, 0, st>>>(col, 3, n, pass, S, d_n);
                LIDAR_CHECK_LAUNCH();
            }
            select_compact_kernel<<<sgrid, kSelThreads, 0, st>>>(col, 3, n, S, buf, d_n);
            LIDAR_CHECK_LAUNCH();
            const int bgrid = sgrid < sm_count() ? sgrid : sm_count();
            for (int pass = 2; pass < 8; ++pass) {
                select_pass_buf_kernel<<<bgrid, kSelThreads, 0, st>>>(buf, pass, S);
                LIDAR_CHECK_LAUNCH();
            }
        } else {
            for (int pass = 0; pass < 8; ++pass) {
                select_pass_kernel<<<sgrid, kSelThreads, 0, st>>>(col, 3, n, pass, S, d_n);
                LIDAR_CHECK_LAUNCH();
            }
        }
        select_next_kernel<<<sgrid, kSelThreads, 0, st>>>(col, 3, n, S, F->kth, F);
        LIDAR_CHECK_LAUNCH();
    }
    // ground split + plane sums + both bboxes
    if ((rc = reset_compaction()) != LIDAR_OK) return rc;
    LIDAR_CUDA_TRY(cudaMemsetAsync(pws + L.off_plane + offsetof(PlaneWs, ticket), 0, sizeof(unsigned), st));
    if (grid > kPlaneMaxBlocks) grid = kPlaneMaxBlocks;
    ground_split_kernel<<<grid, kCmpThreads, 0, st>>>(d_inliers, n, 0.0, 0.0, 0.0, 0.0, d_nonground, d_ng_index,
        &F->n_nonground, F->plane, reinterpret_cast<unsigned long long*>(&F->guard_ground), 0.0,
        reinterpret_cast<unsigned long long*>(pws + L.off_desc), reinterpret_cast<CompactCtrl*>(pws + L.off_ctrl),
        reinterpret_cast<PlaneWs*>(pws + L.off_plane), (int)L.tiles, F);
    LIDAR_CHECK_LAUNCH();
    if (flags & LIDAR_FRONT_SCALER) {
        const long long* d_m = reinterpret_cast<const long long*>(&F->n_nonground);
        if ((rc = moments_f64x3_dev(d_nonground, n, d_m, nullptr, F->t1, rws, st)) != LIDAR_OK) return rc;
        if ((rc = glue(kGlueScMean)) != LIDAR_OK) return rc;
        if ((rc = moments_f64x3_dev(d_nonground, n, d_m, F->sc_mean, F->t2, rws, st)) != LIDAR_OK) return rc;
        if ((rc = glue(kGlueScale)) != LIDAR_OK) return rc;
        standardize_front_kernel<<<ew_grid(n * 3), 256, 0, st>>>(d_nonground, F, d_scaled);
        LIDAR_CHECK_LAUNCH();
        if ((rc = moments_f64x3_dev(d_scaled, n, d_m, nullptr, F->u1, rws, st)) != LIDAR_OK) return rc;
        if ((rc = glue(kGlueXMean)) != LIDAR_OK) return rc;
        if ((rc = moments_f64x3_dev(d_scaled, n, d_m, F->xm, F->u2, rws, st)) != LIDAR_OK) return rc;
        if ((rc = glue(kGlueEps)) != LIDAR_OK) return rc;
    }
    return LIDAR_OK;
}

size_t lidar_preprocess_workspace_bytes(int64_t n) {
    size_t a = pre_layout(n < 0 ? 0 : n).total;
    size_t b = ws_align(sizeof(SelectState)) + ws_align(sizeof(unsigned long long) * (size_t)(n < 0 ? 0 : n));   // + compaction buffer
    return a > b ? a : b;
}

int lidar_sigma_filter(const double* d_points, int64_t n, const double* h_mean3, const double* h_thr3,
                       const double* h_tol3, double zmin, double zden, uint8_t* d_mask, double* d_out_points,
                       double* d_out_colors, int64_t* d_count, uint64_t* d_guard, void* d_ws, size_t ws_bytes,
                       void* stream) {
    LIDAR_REQUIRE(n >= 0 && h_mean3 && h_thr3 && h_tol3 && d_count && d_guard, LIDAR_ERR_INVALID,
                  "lidar_sigma_filter: bad argument");
    LIDAR_REQUIRE(n == 0 || (d_points && d_out_points), LIDAR_ERR_INVALID, "lidar_sigma_filter: NULL points");
    const PreLayout L = pre_layout(n);
    LIDAR_REQUIRE(d_ws && ws_bytes >= L.total, LIDAR_ERR_WORKSPACE, "lidar_sigma_filter: workspace too small (%zu < %zu)",
                  ws_bytes, L.total);
    cudaStream_t st = as_stream(stream);
    char* ws = static_cast<char*>(d_ws);
    LIDAR_CUDA_TRY(cudaMemsetAsync(ws, 0, ws_align(L.off_desc + sizeof(unsigned long long) * L.tiles), st));
    LIDAR_CUDA_TRY(cudaMemsetAsync(d_count, 0, sizeof(int64_t), st));
    LIDAR_CUDA_TRY(cudaMemsetAsync(d_guard, 0, sizeof(uint64_t), st));
    if (n == 0) return LIDAR_OK;
    SigmaParams P;
    for (int c = 0; c < 3; ++c) { P.mean[c] = h_mean3[c]; P.thr[c] = h_thr3[c]; P.tol[c] = h_tol3[c]; }
    P.zmin = zmin; P.zden = zden;
    int grid = sm_count() * 4;
    if ((int64_t)grid > L.tiles) grid = (int)L.tiles;
    sigma_filter_kernel<<<grid, kCmpThreads, 0, st>>>(d_points, n, P, d_mask, d_out_points, d_out_colors, d_count,
        reinterpret_cast<unsigned long long*>(d_guard), reinterpret_cast<unsigned long long*>(ws + L.off_desc),
        reinterpret_cast<CompactCtrl*>(ws + L.off_ctrl), (int)L.tiles);
    LIDAR_CHECK_LAUNCH();
    return LIDAR_OK;
}

int lidar_ground_split(const double* d_points, int64_t n, double z_threshold, const double* h_center3, double tol,
                       double* d_out_points, int32_t* d_out_index, int64_t* d_count, double* d_plane10,
                       uint64_t* d_guard, void* d_ws, size_t ws_bytes, void* stream) {
    LIDAR_REQUIRE(n >= 0 && h_center3 && d_count && d_plane10 && d_guard, LIDAR_ERR_INVALID,
                  "lidar_ground_split: bad argument");
    LIDAR_REQUIRE(n == 0 || (d_points && d_out_points && d_out_index), LIDAR_ERR_INVALID,
                  "lidar_ground_split: NULL points");
    LIDAR_REQUIRE(n < (1ll << 31), LIDAR_ERR_INVALID, "lidar_ground_split: n must be < 2^31");
    const PreLayout L = pre_layout(n);
    LIDAR_REQUIRE(d_ws && ws_bytes >= L.total, LIDAR_ERR_WORKSPACE, "lidar_ground_split: workspace too small");
    cudaStream_t st = as_stream(stream);
    char* ws = static_cast<char*>(d_ws);
    LIDAR_CUDA_TRY(cudaMemsetAsync(ws, 0, ws_align(L.off_desc + sizeof(unsigned long long) * L.tiles), st));
    LIDAR_CUDA_TRY(cudaMemsetAsync(ws + L.off_plane + offsetof(PlaneWs, ticket), 0, sizeof(unsigned), st));
    LIDAR_CUDA_TRY(cudaMemsetAsync(d_count, 0, sizeof(int64_t), st));
    LIDAR_CUDA_TRY(cudaMemsetAsync(d_guard, 0, sizeof(uint64_t), st));
    LIDAR_CUDA_TRY(cudaMemsetAsync(d_plane10, 0, sizeof(double) * 10, st));
    if (n == 0) return LIDAR_OK;
    int grid = sm_count() * 4;
    if ((int64_t)grid > L.tiles) grid = (int)L.tiles;
    if (grid > kPlaneMaxBlocks) grid = kPlaneMaxBlocks;
    ground_split_kernel<<<grid, kCmpThreads, 0, st>>>(d_points, n, z_threshold, h_center3[0], h_center3[1],
        h_center3[2], d_out_points, d_out_index, d_count, d_plane10, reinterpret_cast<unsigned long long*>(d_guard),
        tol, reinterpret_cast<unsigned long long*>(ws + L.off_desc), reinterpret_cast<CompactCtrl*>(ws + L.off_ctrl),
        reinterpret_cast<PlaneWs*>(ws + L.off_plane), (int)L.tiles);
    LIDAR_CHECK_LAUNCH();
    return LIDAR_OK;
}

int lidar_select_kth(const double* d_column, int64_t stride_elems, int64_t n, int64_t k, double* d_out2,
                     void* d_ws, size_t ws_bytes, void* stream) {
    LIDAR_REQUIRE(n > 0 && k >= 0 && k < n, LIDAR_ERR_INVALID, "lidar_select_kth: need 0 <= k < n (k=%lld n=%lld)",
                  (long long)k, (long long)n);
    LIDAR_REQUIRE(d_column && d_out2 && stride_elems > 0, LIDAR_ERR_INVALID, "lidar_select_kth: bad argument");
    LIDAR_REQUIRE(d_ws && ws_bytes >= sizeof(SelectState), LIDAR_ERR_WORKSPACE, "lidar_select_kth: workspace too small");
    cudaStream_t st = as_stream(stream);
    SelectState* S = static_cast<SelectState*>(d_ws);
    select_init_kernel<<<1, 256, 0, st>>>(S, k);
    LIDAR_CHECK_LAUNCH();
    int grid = (int)((n + kSelThreads * 8 - 1) / (kSelThreads * 8));
    const int cap = sm_count() * 8;
    grid = grid < 1 ? 1 : (grid > cap ? cap : grid);
    const size_t buf_off = ws_align(sizeof(SelectState));
    const bool buffered = ws_bytes >= buf_off + sizeof(unsigned long long) * (size_t)n && n >= 65536;
    if (buffered) {
        unsigned long long* buf = reinterpret_cast<unsigned long long*>(static_cast<char*>(d_ws) + buf_off);
        for (int pass = 0; pass < 2; ++pass) {
            select_pass_kernel<<<grid, kSelThreads, 0, st>>>(d_column, stride_elems, n, pass, S);
            LIDAR_CHECK_LAUNCH();
        }
        select_compact_kernel<<<grid, kSelThreads, 0, st>>>(d_column, stride_elems, n, S, buf);
        LIDAR_CHECK_LAUNCH();
        const int bgrid = grid < sm_count() ? grid : sm_count();      // the buffer is small: one CTA per SM at most
        for (int pass = 2; pass < 8; ++pass) {
            select_pass_buf_kernel<<<bgrid, kSelThreads, 0, st>>>(buf, pass, S);
            LIDAR_CHECK_LAUNCH();
        }
    } else {
        for (int pass = 0; pass < 8; ++pass) {
            select_pass_kernel<<<grid, kSelThreads, 0, st>>>(d_column, stride_elems, n, pass, S);
            LIDAR_CHECK_LAUNCH();
        }
    }
    select_next_kernel<<<grid, kSelThreads, 0, st>>>(d_column, stride_elems, n, S, d_out2);
    LIDAR_CHECK_LAUNCH();
    return LIDAR_OK;
}

int lidar_standardize(const double* d_points, int64_t n, const double* h_mean3, const double* h_scale3,
                      double* d_out, void* stream) {
    LIDAR_REQUIRE(n >= 0 && h_mean3 && h_scale3, LIDAR_ERR_INVALID, "lidar_standardize: bad argument");
    if (n == 0) return LIDAR_OK;
    LIDAR_REQUIRE(d_points && d_out, LIDAR_ERR_INVALID, "lidar_standardize: NULL points");
    standardize_kernel<<<ew_grid(n * 3), 256, 0, as_stream(stream)>>>(d_points, n * 3, h_mean3[0], h_mean3[1],
        h_mean3[2], h_scale3[0], h_scale3[1], h_scale3[2], d_out);
    LIDAR_CHECK_LAUNCH();
    return LIDAR_OK;
}

int lidar_gather_rows(const void* d_src, int64_t n_rows, int row_bytes, const int64_t* d_index, int64_t k,
                      void* d_dst, void* stream) {
    LIDAR_REQUIRE(k >= 0 && n_rows >= 0 && row_bytes > 0 && row_bytes % 4 == 0, LIDAR_ERR_INVALID,
                  "lidar_gather_rows: row_bytes must be a positive multiple of 4");
    if (k == 0) return LIDAR_OK;
    LIDAR_REQUIRE(d_src && d_index && d_dst, LIDAR_ERR_INVALID, "lidar_gather_rows: NULL argument");
    gather_rows_kernel<<<ew_grid(k * (row_bytes / 4)), 256, 0, as_stream(stream)>>>(
        static_cast<const uint32_t*>(d_src), row_bytes / 4, reinterpret_cast<const long long*>(d_index), k,
        static_cast<uint32_t*>(d_dst));
    LIDAR_CHECK_LAUNCH();
    return LIDAR_OK;
}

int lidar_scatter_labels(const int32_t* d_labels, const int32_t* d_index, int64_t m, int64_t* d_full, int64_t n,
                         void* stream) {
    LIDAR_REQUIRE(m >= 0 && n >= 0, LIDAR_ERR_INVALID, "lidar_scatter_labels: bad sizes");
    if (n == 0) return LIDAR_OK;
    LIDAR_REQUIRE(d_full, LIDAR_ERR_INVALID, "lidar_scatter_labels: NULL output");
    cudaStream_t st = as_stream(stream);
    fill_i64_kernel<<<ew_grid(n), 256, 0, st>>>(reinterpret_cast<long long*>(d_full), n, -1ll);
    LIDAR_CHECK_LAUNCH();
    if (m > 0) {
        LIDAR_REQUIRE(d_labels && d_index, LIDAR_ERR_INVALID, "lidar_scatter_labels: NULL labels");
        scatter_labels_kernel<<<ew_grid(m), 256, 0, st>>>(d_labels, d_index, m, reinterpret_cast<long long*>(d_full));
        LIDAR_CHECK_LAUNCH();
    }
    return LIDAR_OK;
}

}  // extern "C"
