// K11 / K12 — PointNet++-style set abstraction: farthest point sampling, ball query, grouping.
//
// NEW ops: the reference names a classifier in a design note (windows_design.md:65) but contains no
// FPS / ball query / grouping code at all, so there is no call site to cite; the
// contracts are SURVEY.md Appendix B.4-B.6 (published PointNet++ CUDA-op semantics).
//
//   lidar_fps          one thread-block CLUSTER per cloud.  Every CTA of the cluster keeps its slice of
//                      the cloud (x, y, z and the running min distance) in registers; per iteration a
//                      warp-shuffle argmax, a CTA argmax through shared memory, and a cluster argmax
//                      through DISTRIBUTED shared memory (each CTA stores its candidate into every
//                      peer's slot array with st.shared::cluster, one barrier.cluster per iteration,
//                      slots double-buffered).  Latency-bound by design: M dependent steps.
//                      fp32 ((dx*dx + dy*dy) + dz*dz, no FMA), lowest index wins ties -> bit-exact.
//   lidar_ball_query   tiles of the cloud are staged in shared memory with the TMA bulk-copy engine
//                      (cp.async.bulk + mbarrier, double buffered); each warp owns 4 centres and walks
//                      the tile 32 points at a time, ballot + popc keep the hits in ascending index
//                      order; strict d2 < r2 in fp32 with the same expression as FPS -> bit-exact.
//   lidar_group_points gather + centre subtraction (exact fp32), features gathered unchanged.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace lidar {

// ------------------------------------------------------------------------------------------------
// FPS
// ------------------------------------------------------------------------------------------------
constexpr int kFpsThreads = 256;
constexpr int kFpsPpt = 8;          // points per thread held in registers (a CONTIGUOUS run of the cloud)
constexpr int kFpsMaxCluster = 8;
constexpr int kFpsWarps = kFpsThreads / 32;

__device__ __forceinline__ bool cand_better(float v, int i, float bv, int bi) {
    return v > bv || (v == bv && i < bi);
}
__device__ __forceinline__ unsigned fps_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned fps_map_peer(unsigned local_addr, unsigned peer) {
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(peer));
    return r;
}
// remote store + transaction-count signal on the peer's mbarrier: data and "it has arrived" travel together,
// the receiver wakes from mbarrier.try_wait ~60 cycles after the last byte (a barrier.cluster round costs ~400-500)
__device__ __forceinline__ void fps_send_cand(unsigned slot_addr, unsigned bar_addr, unsigned val, float x, float y, float z) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
                 ::"r"(slot_addr), "r"(val), "r"(__float_as_uint(x)), "r"(__float_as_uint(y)), "r"(__float_as_uint(z)),
                 "r"(bar_addr) : "memory");
}
// (default .acquire.cta: shared memory is not cached in L1, and a cluster-scope acquire would add a CCTL.IVALL
// to every iteration; the phase completion itself orders the peers' st.async data before this read)
__device__ __forceinline__ void fps_mbar_wait(unsigned bar_addr, unsigned parity) {
    unsigned done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar_addr), "r"(parity) : "memory");
    } while (!done);
}

// index of the first of eight values equal to their maximum (all lanes compute it redundantly: a depth-3 max
// tree and an equality mask cost ~60 cycles, a REDUX + ballot + shuffle round ~130)
__device__ __forceinline__ int fps_first_max8(const unsigned (&v)[8], unsigned& mx) {
    const unsigned a = max(v[0], v[1]), b = max(v[2], v[3]), c = max(v[4], v[5]), d = max(v[6], v[7]);
    mx = max(max(a, b), max(c, d));
    unsigned mask = 0u;
#pragma unroll
    for (int k = 0; k < 8; ++k) mask |= (v[k] == mx ? 1u : 0u) << k;
    return __ffs(mask) - 1;
}

// One thread-block cluster per cloud.
//   * thread t of CTA r owns the contiguous indices [(r*256 + t)*8, +8): position order == index order at every
//     level (lane, warp, CTA), so "lowest index among equals" is "lowest position": no index reduction, and the
//     index does not even travel -- the CTA that owns the winner writes it out
//   * distances are >= 0, so their fp32 bit patterns order like the values: the warp argmax is ONE REDUX.MAX
//     plus a ballot (lowest lane holding the maximum writes the warp's entry)
//   * CTA argmax: warp 0 folds the 8 warp entries and sends the CTA's 16-byte candidate {value bits, x, y, z} to
//     EVERY CTA of the cluster (one st.async + complete_tx on the receiver's mbarrier; slots and barriers double
//     buffered) -- one DSMEM transit and one mbarrier wake-up per iteration instead of a barrier.cluster round
//   * 25 KB of shared memory per CTA (its own slice, for the coordinates of its candidate), so the 16 clusters of
//     a 16-cloud batch are co-resident (with the whole cloud per CTA only 15 clusters fit: two waves)
template <int kCluster>
__global__ void __launch_bounds__(kFpsThreads)
fps_cluster_kernel(const float* __restrict__ xyz, int n, int m, int* __restrict__ out) {
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned rank = cluster.block_rank();
    const int cloud = blockIdx.x / kCluster;
    const float* P = xyz + (size_t)cloud * n * 3;
    int* O = out + (size_t)cloud * m;
    constexpr int per_cta = kFpsThreads * kFpsPpt;
    static_assert(kFpsWarps == 8 && kFpsMaxCluster == 8, "fps_first_max8 folds exactly eight entries");

    __shared__ float s_x[per_cta], s_y[per_cta], s_z[per_cta];   // this CTA's slice, SoA
    __shared__ __align__(16) uint4 s_slots[2][kFpsMaxCluster];   // one candidate per CTA of the cluster (DSMEM)
    __shared__ __align__(16) unsigned s_wval[2][kFpsWarps];      // per-warp candidates of this CTA
    __shared__ __align__(16) unsigned s_widx[2][kFpsWarps];
    __shared__ __align__(8) unsigned long long s_mbar[2];

    const int tid = threadIdx.x;
    const unsigned lane = lane_id();
    const int warp = tid >> 5;
    const int slice0 = rank * per_cta;
    for (int l = tid; l < per_cta; l += kFpsThreads) {
        const int i = slice0 + l;
        const bool live = i < n;
        s_x[l] = live ? P[3 * i] : 0.f;
        s_y[l] = live ? P[3 * i + 1] : 0.f;
        s_z[l] = live ? P[3 * i + 2] : 0.f;
    }
    // slots of absent CTAs (kCluster < 8) stay at value 0: they sit after the real ones and never win a tie
    if (tid < 2 * kFpsMaxCluster) (&s_slots[0][0])[tid] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    float px[kFpsPpt], py[kFpsPpt], pz[kFpsPpt], md[kFpsPpt];
    const int base_idx = slice0 + tid * kFpsPpt;
#pragma unroll
    for (int j = 0; j < kFpsPpt; ++j) {
        const int l = tid * kFpsPpt + j;
        px[j] = s_x[l]; py[j] = s_y[l]; pz[j] = s_z[l];
        md[j] = slice0 + l < n ? 1e10f : -1.f;   // padding never wins (real distances are >= 0)
    }
    float sx = P[0], sy = P[1], sz = P[2];   // idx[0] = 0
    if (rank == 0 && tid == 0) O[0] = 0;
    const unsigned bar0 = fps_smem_u32(&s_mbar[0]), bar1 = fps_smem_u32(&s_mbar[1]);
    constexpr unsigned kTxBytes = kCluster * 16u;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0), "r"(kTxBytes) : "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar1), "r"(kTxBytes) : "memory");
    }
    cluster.sync();   // every peer's barriers and slots exist before the first remote store
    // lane r of warp 0 talks to CTA r: this CTA's slot there, and that CTA's barriers
    const unsigned peer = lane < (unsigned)kCluster ? lane : 0u;
    const unsigned peer_slot0 = fps_map_peer(fps_smem_u32(&s_slots[0][rank]), peer);
    const unsigned peer_slot1 = fps_map_peer(fps_smem_u32(&s_slots[1][rank]), peer);
    const unsigned peer_bar0 = fps_map_peer(bar0, peer), peer_bar1 = fps_map_peer(bar1, peer);

    for (int it = 1; it < m; ++it) {
        const int par = it & 1;
        const unsigned phase = (unsigned)((it - 1) >> 1) & 1u;   // k-th use of barrier `par`
        float bv = -1.f;
        int bi = 0x7fffffff;
#pragma unroll
        for (int j = 0; j < kFpsPpt; ++j) {
            const float dx = __fsub_rn(px[j], sx), dy = __fsub_rn(py[j], sy), dz = __fsub_rn(pz[j], sz);
            const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            const float v = md[j] < 0.f ? md[j] : fminf(md[j], d);
            md[j] = v;
            if (v > bv) { bv = v; bi = base_idx + j; }    // ascending j: strict compare keeps the lowest index
        }
        // warp argmax: one REDUX over the value bits, the lowest lane holding the maximum writes the entry
        const unsigned vb = bv < 0.f ? 0u : __float_as_uint(bv);
        const unsigned wm = __reduce_max_sync(0xffffffffu, vb);
        const unsigned holders = __ballot_sync(0xffffffffu, vb == wm);
        if ((holders & lanemask_lt()) == 0u && vb == wm) { s_wval[par][warp] = wm; s_widx[par][warp] = (unsigned)bi; }
        __syncthreads();
        int my_ci = 0;
        if (warp == 0) {
            const uint4 a = *reinterpret_cast<const uint4*>(&s_wval[par][0]);
            const uint4 b = *reinterpret_cast<const uint4*>(&s_wval[par][4]);
            const unsigned v8[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
            unsigned cm;
            const int pos = fps_first_max8(v8, cm);
            my_ci = (int)s_widx[par][pos];
            const int l = my_ci != 0x7fffffff ? my_ci - slice0 : 0;   // an all-padding CTA sends value 0
            if (lane < (unsigned)kCluster)
                fps_send_cand(par ? peer_slot1 : peer_slot0, par ? peer_bar1 : peer_bar0, cm, s_x[l], s_y[l], s_z[l]);
        }
        fps_mbar_wait(par ? bar1 : bar0, phase);
        // every thread folds the candidates of the cluster's CTAs (slot order == index order)
        unsigned g8[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) g8[r] = s_slots[par][r].x;
        unsigned gm;
        const int wp = fps_first_max8(g8, gm);
        const uint4 w = s_slots[par][wp];
        sx = __uint_as_float(w.y); sy = __uint_as_float(w.z); sz = __uint_as_float(w.w);
        if (tid == 0) {
            if (wp == (int)rank) O[it] = my_ci;     // the owner of the winner writes its index
            // re-arm this barrier for its next use (iteration it + 2): peers cannot send for it + 2 before they
            // have received this CTA's candidate of it + 1, which leaves after this point
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                         ::"r"(par ? bar1 : bar0), "r"(kTxBytes) : "memory");
        }
    }
    cluster.sync();   // nobody exits while a peer may still write into its slots
}

// generic fallback for clouds larger than one cluster's register capacity: one CTA per cloud, points
// re-read from global / L2 every iteration (used only when n > 8 * 256 * 8 = 16384)
__global__ void __launch_bounds__(1024)
fps_block_kernel(const float* __restrict__ xyz, int n, int m, float* __restrict__ mind, int* __restrict__ out) {
    const int cloud = blockIdx.x;
    const float* P = xyz + (size_t)cloud * n * 3;
    float* MD = mind + (size_t)cloud * n;
    int* O = out + (size_t)cloud * m;
    __shared__ float s_v[32];
    __shared__ int s_i[32];
    __shared__ int s_cur;
    for (int i = threadIdx.x; i < n; i += blockDim.x) MD[i] = 1e10f;
    if (threadIdx.x == 0) { O[0] = 0; s_cur = 0; }
    __syncthreads();
    for (int it = 1; it < m; ++it) {
        const int cur = s_cur;
        const float sx = P[3 * cur], sy = P[3 * cur + 1], sz = P[3 * cur + 2];
        float bv = -2.f;
        int bi = 0x7fffffff;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const float dx = __fsub_rn(P[3 * i], sx), dy = __fsub_rn(P[3 * i + 1], sy), dz = __fsub_rn(P[3 * i + 2], sz);
            const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            const float v = fminf(MD[i], d);
            MD[i] = v;
            if (cand_better(v, i, bv, bi)) { bv = v; bi = i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (cand_better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
        }
        if (lane_id() == 0) { s_v[threadIdx.x >> 5] = bv; s_i[threadIdx.x >> 5] = bi; }
        __syncthreads();
        if (threadIdx.x < 32) {
            bv = threadIdx.x < (blockDim.x >> 5) ? s_v[threadIdx.x] : -2.f;
            bi = threadIdx.x < (blockDim.x >> 5) ? s_i[threadIdx.x] : 0x7fffffff;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (cand_better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
            }
            if (threadIdx.x == 0) { s_cur = bi; O[it] = bi; }
        }
        __syncthreads();
    }
}

template <int kCluster>
static cudaError_t launch_fps(const float* xyz, int b, int n, int m, int* out, cudaStream_t st) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(b * kCluster);
    cfg.blockDim = dim3(kFpsThreads);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kCluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, fps_cluster_kernel<kCluster>, xyz, n, m, out);
}

// ------------------------------------------------------------------------------------------------
// ball query
// ------------------------------------------------------------------------------------------------
constexpr int kBqThreads = 256;
constexpr int kBqCentresPerWarp = 4;
constexpr int kBqCentresPerCta = (kBqThreads / 32) * kBqCentresPerWarp;   // 32
constexpr int kBqTile = 1920;                                             // points per staged tile (22.5 KB)

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(kBqThreads)
ball_query_kernel(const float* __restrict__ xyz, const float* __restrict__ new_xyz, int n, int m, float r2, int k,
                  int use_tma, int* __restrict__ idx_out) {
    __shared__ __align__(128) float s_tile[2][kBqTile * 3];
    __shared__ __align__(8) unsigned long long s_bar[2];
    const int ctas_per_cloud = (m + kBqCentresPerCta - 1) / kBqCentresPerCta;
    const int cloud = blockIdx.x / ctas_per_cloud;
    const int c0 = (blockIdx.x % ctas_per_cloud) * kBqCentresPerCta;
    const float* P = xyz + (size_t)cloud * n * 3;
    const int warp = threadIdx.x >> 5;
    const unsigned lane = lane_id();
    const int n_tiles = (n + kBqTile - 1) / kBqTile;

    float cx[kBqCentresPerWarp], cy[kBqCentresPerWarp], cz[kBqCentresPerWarp];
    int cnt[kBqCentresPerWarp];
    int* outp[kBqCentresPerWarp];
#pragma unroll
    for (int q = 0; q < kBqCentresPerWarp; ++q) {
        const int c = c0 + warp * kBqCentresPerWarp + q;
        const bool live = c < m;
        const float* C = new_xyz + ((size_t)cloud * m + (live ? c : 0)) * 3;
        cx[q] = C[0]; cy[q] = C[1]; cz[q] = C[2];
        cnt[q] = live ? 0 : k;   // dead slots are "full": they never record anything
        outp[q] = idx_out + ((size_t)cloud * m + (live ? c : 0)) * k;
        if (live)
            for (int j = lane; j < k; j += 32) outp[q][j] = 0;   // PointNet++: no hit at all leaves zeros
    }
    if (use_tma && threadIdx.x == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto tile_points = [&](int t) { return n - t * kBqTile < kBqTile ? n - t * kBqTile : kBqTile; };
    auto issue = [&](int t) {
        const int pts = tile_points(t);
        if (use_tma) {
            if (threadIdx.x == 0) {
                const unsigned bytes = (unsigned)pts * 12u;
                mbar_expect_tx(&s_bar[t & 1], bytes);
                tma_bulk_g2s(s_tile[t & 1], P + (size_t)t * kBqTile * 3, bytes, &s_bar[t & 1]);
            }
        } else {
            for (int j = threadIdx.x; j < pts * 3; j += kBqThreads) s_tile[t & 1][j] = P[(size_t)t * kBqTile * 3 + j];
        }
    };
    issue(0);
    for (int t = 0; t < n_tiles; ++t) {
        if (t + 1 < n_tiles) issue(t + 1);          // prefetch the next tile into the other buffer
        if (use_tma) mbar_wait(&s_bar[t & 1], (t >> 1) & 1); else __syncthreads();
        const int pts = tile_points(t);
        const float* T = s_tile[t & 1];
        const bool warp_done = cnt[0] >= k && cnt[1] >= k && cnt[2] >= k && cnt[3] >= k;
        if (!warp_done) {
            for (int j0 = 0; j0 < pts; j0 += 32) {
                const int j = j0 + (int)lane;
                const bool in = j < pts;
                const float x = in ? T[3 * j] : 0.f, y = in ? T[3 * j + 1] : 0.f, z = in ? T[3 * j + 2] : 0.f;
#pragma unroll
                for (int q = 0; q < kBqCentresPerWarp; ++q) {
                    if (cnt[q] >= k) continue;      // warp-uniform
                    const float dx = __fsub_rn(cx[q], x), dy = __fsub_rn(cy[q], y), dz = __fsub_rn(cz[q], z);
                    const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
                    const unsigned hits = __ballot_sync(0xffffffffu, in && d2 < r2);
                    if (hits) {
                        const int gidx = t * kBqTile + j;
                        if (cnt[q] == 0) {
                            // first hit pre-fills every slot
                            const int first = t * kBqTile + j0 + (__ffs(hits) - 1);
                            for (int s = lane; s < k; s += 32) outp[q][s] = first;
                            __syncwarp();
                        }
                        const int pos = cnt[q] + __popc(hits & lanemask_lt());
                        if (((hits >> lane) & 1u) && pos < k) outp[q][pos] = gidx;
                        cnt[q] += __popc(hits);
                    }
                }
                if (cnt[0] >= k && cnt[1] >= k && cnt[2] >= k && cnt[3] >= k) break;
            }
        }
        __syncthreads();   // everyone is done with buffer t&1 before tile t+2 overwrites it
    }
}

// ------------------------------------------------------------------------------------------------
// grouping
// ------------------------------------------------------------------------------------------------
__global__ void group_points_kernel(const float* __restrict__ xyz, const float* __restrict__ feats,
                                    const int* __restrict__ idx, const float* __restrict__ new_xyz, int b, int n,
                                    int m, int k, int c_feat, float* __restrict__ out) {
    const int64_t total = (int64_t)b * (3 + c_feat) * m * k;
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += step) {
        const int j = (int)(t % k);
        int64_t r = t / k;
        const int mm = (int)(r % m);
        r /= m;
        const int c = (int)(r % (3 + c_feat));
        const int bb = (int)(r / (3 + c_feat));
        const int src = idx[((size_t)bb * m + mm) * k + j];
        float v;
        if (c < 3) v = __fsub_rn(xyz[((size_t)bb * n + src) * 3 + c], new_xyz[((size_t)bb * m + mm) * 3 + c]);
        else v = feats[((size_t)bb * c_feat + (c - 3)) * n + src];
        out[t] = v;
    }
}

__global__ void gather_points_kernel(const float* __restrict__ xyz, const int* __restrict__ idx, int b, int n, int m,
                                     float* __restrict__ out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= b * m) return;
    const int bb = t / m;
    const int src = idx[t];
    const float* p = xyz + ((size_t)bb * n + src) * 3;
    out[3 * (size_t)t] = p[0]; out[3 * (size_t)t + 1] = p[1]; out[3 * (size_t)t + 2] = p[2];
}

}  // namespace lidar

using namespace lidar;

extern "C" {

size_t lidar_fps_workspace_bytes(int b, int n) {
    if (b < 0 || n < 0) return 0;
    return n > kFpsThreads * kFpsPpt * kFpsMaxCluster ? ws_align(sizeof(float) * (size_t)b * n) : 256;
}

int lidar_fps(const float* d_xyz, int b, int n, int m, int32_t* d_idx, void* d_ws, size_t ws_bytes, void* stream) {
    LIDAR_REQUIRE(b >= 0 && n >= 1 && m >= 1 && m <= n, LIDAR_ERR_INVALID, "lidar_fps: need 1 <= m <= n (b=%d n=%d m=%d)", b, n, m);
    if (b == 0) return LIDAR_OK;
    LIDAR_REQUIRE(d_xyz && d_idx, LIDAR_ERR_INVALID, "lidar_fps: NULL argument");
    cudaStream_t st = as_stream(stream);
    const int per_cta = kFpsThreads * kFpsPpt;
    if (n <= per_cta) LIDAR_CUDA_TRY(launch_fps<1>(d_xyz, b, n, m, d_idx, st));
    else if (n <= 2 * per_cta) LIDAR_CUDA_TRY(launch_fps<2>(d_xyz, b, n, m, d_idx, st));
    else if (n <= 4 * per_cta) LIDAR_CUDA_TRY(launch_fps<4>(d_xyz, b, n, m, d_idx, st));
    else if (n <= 8 * per_cta) LIDAR_CUDA_TRY(launch_fps<8>(d_xyz, b, n, m, d_idx, st));
    else {
        const size_t need = sizeof(float) * (size_t)b * n;
        LIDAR_REQUIRE(d_ws && ws_bytes >= need, LIDAR_ERR_WORKSPACE, "lidar_fps: workspace too small (%zu < %zu)", ws_bytes, need);
        fps_block_kernel<<<b, 1024, 0, st>>>(d_xyz, n, m, static_cast<float*>(d_ws), d_idx);
    }
    LIDAR_CHECK_LAUNCH();
    return LIDAR_OK;
}

int lidar_ball_query(const float* d_xyz, const float* d_new_xyz, int b, int n, int m, float radius, int k,
                     int32_t* d_idx, void* stream) {
    LIDAR_REQUIRE(b >= 0 && n >= 1 && m >= 1 && k >= 1 && radius >= 0.f, LIDAR_ERR_INVALID, "lidar_ball_query: bad sizes");
    if (b == 0) return LIDAR_OK;
    LIDAR_REQUIRE(d_xyz && d_new_xyz && d_idx, LIDAR_ERR_INVALID, "lidar_ball_query: NULL argument");
    // the bulk-copy engine needs 16-byte aligned sources and sizes: clouds of 4k points from a 16 B aligned base
    const int use_tma = (n % 4 == 0) && ((reinterpret_cast<uintptr_t>(d_xyz) & 15u) == 0);
    const int ctas_per_cloud = (m + kBqCentresPerCta - 1) / kBqCentresPerCta;
    ball_query_kernel<<<b * ctas_per_cloud, kBqThreads, 0, as_stream(stream)>>>(d_xyz, d_new_xyz, n, m,
                                                                                 radius * radius, k, use_tma, d_idx);
    LIDAR_CHECK_LAUNCH();
    return LIDAR_OK;
}

int lidar_group_points(const float* d_xyz, const float* d_feats, const int32_t* d_idx, const float* d_new_xyz, int b,
                       int n, int m, int k, int c_feat, float* d_out, void* stream) {
    LIDAR_REQUIRE(b >= 0 && n >= 1 && m >= 1 && k >= 1 && c_feat >= 0, LIDAR_ERR_INVALID, "lidar_group_points: bad sizes");
    if (b == 0) return LIDAR_OK;
    LIDAR_REQUIRE(d_xyz && d_idx && d_new_xyz && d_out && (c_feat == 0 || d_feats), LIDAR_ERR_INVALID,
                  "lidar_group_points: NULL argument");
    const int64_t total = (int64_t)b * (3 + c_feat) * m * k;
    int64_t want = (total + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 16;
    group_points_kernel<<<(int)(want < cap ? want : cap), 256, 0, as_stream(stream)>>>(d_xyz, d_feats, d_idx, d_new_xyz,
                                                                                       b, n, m, k, c_feat, d_out);
    LIDAR_CHECK_LAUNCH();
    return LIDAR_OK;
}

int lidar_gather_points(const float* d_xyz, const int32_t* d_idx, int b, int n, int m, float* d_out, void* stream) {
    LIDAR_REQUIRE(b >= 0 && n >= 1 && m >= 1, LIDAR_ERR_INVALID, "lidar_gather_points: bad sizes");
    if (b == 0) return LIDAR_OK;
    LIDAR_REQUIRE(d_xyz && d_idx && d_out, LIDAR_ERR_INVALID, "lidar_gather_points: NULL argument");
    gather_points_kernel<<<(b * m + 255) / 256, 256, 0, as_stream(stream)>>>(d_xyz, d_idx, b, n, m, d_out);
    LIDAR_CHECK_LAUNCH();
    return LIDAR_OK;
}

}  // extern "C"
