// K13 (tensor-core path) — fused gather + shared MLP 3-64-64-128 + max-pool over k = 32 on the
// 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM).  NEW op, SURVEY.md Appendix B.7.
//
// One CTA tile = 128 rows = 4 centres x 32 neighbours, so TMEM lane r / row r / lane quarter w line up with
// (neighbour r % 32 of centre w): the max-pool over the neighbours stays inside a warp, no shared memory.
// 256 threads: warps w and w + 4 share lane quarter w (a warp may only touch TMEM lanes 32 (w % 4) ...) and split
// the CHANNELS of every phase between them -- 16 resident warps per SM instead of 8 (the kernel is bound by issue
// slots: 2 warps per scheduler issued 31 % of the cycles, profiles/r2_ncu_mlp_v1_raw.csv).
//   layer 1 (K = 3)        CUDA cores, straight from the gathered (xyz[idx] - centre)
//   layer 2 (128x64x64)    tcgen05.mma kind::f16 (bf16 in, fp32 accumulate in TMEM columns 0..63)
//   layer 3 (128x128x64)   tcgen05.mma kind::f16 (TMEM columns 64..191)
// Precision: rtol 1e-3 is not reachable with plain bf16 (or tf32) operands, so every operand is split
// x = hi + lo with hi = bf16(x), lo = bf16(x - hi) and each GEMM is issued as hi*hi + hi*lo + lo*hi
// (the dropped lo*lo term is ~2^-16 relative) — three MMAs per K step, still tensor-bound trivial.
// Operands are K-major, un-swizzled "interleaved" core matrices (8 rows x 16 B), built directly by the
// producing threads with 16-byte shared stores:
//     offset(row, k) = (row/8)*1024 + (k/8)*128 + (row%8)*16 + (k%8)*2      [SBO = 1024 B, LBO = 128 B]
// Descriptor / instruction-descriptor bit layouts follow the PTX ISA tcgen05 matrix-descriptor tables.
#include <cuda_bf16.h>

#include "common.cuh"

namespace lidar {

constexpr int kTcThreads = 256;     // two warps per TMEM lane quarter: each takes half of the channels of its 32 rows
constexpr int kTcRows = 128;
constexpr int kTcK = 32;            // neighbours per centre
constexpr int kC1 = 64, kC2 = 64, kC3 = 128;
constexpr int kTmemCols = 256;      // 64 (layer 2) + 128 (layer 3) rounded up to a power of two

// shared-memory carve-up (bytes)
constexpr int kOffAh = 0;                       // activations hi  [128 x 64] bf16
constexpr int kOffAl = kOffAh + 128 * 64 * 2;   // activations lo
constexpr int kOffW2h = kOffAl + 128 * 64 * 2;  // W2 hi [64 x 64]
constexpr int kOffW2l = kOffW2h + 64 * 64 * 2;
constexpr int kOffW3h = kOffW2l + 64 * 64 * 2;  // W3 hi [128 x 64]
constexpr int kOffW3l = kOffW3h + 128 * 64 * 2;
constexpr int kOffW1 = kOffW3l + 128 * 64 * 2;  // fp32 [64][3]
constexpr int kOffB1 = kOffW1 + 64 * 3 * 4;
constexpr int kOffB2 = kOffB1 + 64 * 4;
constexpr int kOffB3 = kOffB2 + 64 * 4;
constexpr int kOffBar = kOffB3 + 128 * 4;       // mbarrier (8 B) + TMEM base (4 B)
constexpr int kTcSmem = kOffBar + 16;

__device__ __forceinline__ unsigned s_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// K-major, SWIZZLE_NONE shared-memory matrix descriptor (sm_100 version bit set)
__device__ __forceinline__ unsigned long long umma_desc(unsigned smem_byte_addr) {
    const unsigned long long lbo = 128 >> 4, sbo = 1024 >> 4;
    return (unsigned long long)((smem_byte_addr >> 4) & 0x3FFF) | (lbo << 16) | (sbo << 32) | (1ull << 46);
}
// instruction descriptor: D = F32, A = B = BF16, both K-major, M x N
__device__ __forceinline__ unsigned umma_idesc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(unsigned tmem_d, unsigned long long a, unsigned long long b, unsigned idesc,
                                         unsigned accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(a), "l"(b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned long long* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s_addr(bar)) : "memory");
}
__device__ __forceinline__ void bar_wait(unsigned long long* bar, unsigned parity) {
    unsigned done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(s_addr(bar)), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tmem_ld16(unsigned taddr, float (&v)[16]) {
    unsigned r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_ld16_nowait(unsigned taddr, unsigned (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_16x256b_x4(unsigned taddr, unsigned (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}

// Packed fp32 arithmetic of sm_100 (add / sub / fma .f32x2 on a 64-bit register pair, SASS FADD2 / FFMA2): two IEEE
// operations per issue slot, each rounded exactly like the scalar instruction -- this kernel is bound by issue slots.
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long sub2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

// x = hi + lo with hi = bf16(x), lo = bf16(x - hi), two values per instruction: cvt.rn.bf16x2.f32 (F2FP, a
// full-rate ALU op; the scalar F2F conversion runs at a quarter of that and was 21 % of the kernel's stall samples)
__device__ __forceinline__ void split_pair(float a, float b, unsigned& hi, unsigned& lo) {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));           // low half <- a, high half <- b
    float ra, rb;
    unpack2(sub2(pack2(a, b), pack2(__uint_as_float(hi << 16), __uint_as_float(hi & 0xffff0000u))), ra, rb);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(rb), "f"(ra));
}
// split 8 fp32 values into bf16 hi / lo and store them as one 16-byte K chunk each
__device__ __forceinline__ void store_chunk_at(unsigned char* hi_base, unsigned char* lo_base, int row, int kc, const float (&x)[8]) {
    unsigned hi[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) split_pair(x[2 * i], x[2 * i + 1], hi[i], lo[i]);
    const int off = (row >> 3) * 1024 + kc * 128 + (row & 7) * 16;
    *reinterpret_cast<uint4*>(hi_base + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(lo_base + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}
__device__ __forceinline__ void store_chunk(unsigned char* smem, int row, int kc, const float (&x)[8]) {
    store_chunk_at(smem + kOffAh, smem + kOffAl, row, kc, x);
}

// stage an fp32 [rows x 64] row-major weight matrix as bf16 hi / lo core matrices: one 16-byte K chunk
// (8 consecutive k of one row = two float4 loads) per thread and step
__device__ __forceinline__ void stage_weights(const float* __restrict__ W, int rows, unsigned char* hi_base,
                                              unsigned char* lo_base) {
    for (int e = threadIdx.x; e < rows * 8; e += kTcThreads) {
        const int r = e >> 3, kc = e & 7;
        const float4 a = __ldg(reinterpret_cast<const float4*>(W + (size_t)r * 64 + kc * 8));
        const float4 b = __ldg(reinterpret_cast<const float4*>(W + (size_t)r * 64 + kc * 8) + 1);
        const float x[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        store_chunk_at(hi_base, lo_base, r, kc, x);
    }
}

__global__ void __launch_bounds__(kTcThreads, 2)
shared_mlp_tc_kernel(const float* __restrict__ xyz, const int* __restrict__ idx, const float* __restrict__ new_xyz,
                     int n, int m, int n_centres, const float* __restrict__ W1, const float* __restrict__ B1,
                     const float* __restrict__ W2, const float* __restrict__ B2, const float* __restrict__ W3,
                     const float* __restrict__ B3, float* __restrict__ out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float* sW1 = reinterpret_cast<float*>(smem + kOffW1);
    float* sB2 = reinterpret_cast<float*>(smem + kOffB2);
    float* sB3 = reinterpret_cast<float*>(smem + kOffB3);
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(smem + kOffBar);
    unsigned* tmem_slot = reinterpret_cast<unsigned*>(smem + kOffBar + 8);
    unsigned* arrive_cnt = reinterpret_cast<unsigned*>(smem + kOffBar + 12);
    const int warp = threadIdx.x >> 5;
    const int qwarp = warp & 3;               // TMEM lane quarter = centre of the tile
    const int half = threadIdx.x >> 7;        // which half of the channels this thread works on
    const unsigned lane = lane_id();

    // ---- one-time setup: weights, barrier, TMEM ------------------------------------------------
    stage_weights(W2, kC2, smem + kOffW2h, smem + kOffW2l);
    stage_weights(W3, kC3, smem + kOffW3h, smem + kOffW3l);
    // layer 1 as channel PAIRS for the packed FMA: {w0(o), w0(o+1), w1(o), w1(o+1), w2(o), w2(o+1), b(o), b(o+1)}, o = 2p
    // (two 16-byte shared loads per pair instead of eight scalar ones; kOffW1 .. kOffB2 is one 1 KB block)
    for (int i = threadIdx.x; i < kC1 * 4; i += kTcThreads) {
        const int pr = i >> 3, e = i & 7, o = 2 * pr + (e & 1), w = e >> 1;
        sW1[i] = w < 3 ? W1[o * 3 + w] : B1[o];
    }
    for (int i = threadIdx.x; i < kC2; i += kTcThreads) sB2[i] = B2[i];
    for (int i = threadIdx.x; i < kC3; i += kTcThreads) sB3[i] = B3[i];
    if (threadIdx.x == 0) {
        *arrive_cnt = 0u;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s_addr(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_addr(tmem_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem_base = *tmem_slot;
    const unsigned tmem_lane = tmem_base + ((unsigned)(qwarp * 32) << 16);   // this warp's 32 TMEM lanes
    const unsigned sbase = s_addr(smem);
    const unsigned idesc2 = umma_idesc(kTcRows, kC2), idesc3 = umma_idesc(kTcRows, kC3);
    unsigned phase = 0;

    const int n_tiles = (n_centres + 3) / 4;
    // gathered (neighbour - centre) of a tile = two DEPENDENT global loads (neighbour index, then the point).  Both
    // are software-pipelined: during tile t the index of tile t+2 and the point of tile t+1 (whose index arrived one
    // tile ago) are in flight, so neither latency is exposed (the source-level profile had 10 % of the stall samples
    // on the address computation that waits for the index).
    auto load_index = [&](int t) -> int {
        const int c_ = t * 4 + qwarp;
        return (t < n_tiles && c_ < n_centres) ? __ldg(idx + (size_t)c_ * kTcK + lane) : -1;
    };
    // centre -> (batch, centre inside the batch) without a division per tile: the centres a warp visits advance by a fixed
    // step (the emulated 32-bit divisions were 6 % of the kernel's instructions)
    const int cstep = 4 * (int)gridDim.x;
    auto advance = [&](int& b_, int& mm_) {
        mm_ += cstep;
        if (cstep < 8 * m) {
            while (mm_ >= m) { mm_ -= m; ++b_; }
        } else {
            b_ += mm_ / m;
            mm_ %= m;
        }
    };
    const int c_first = (int)blockIdx.x * 4 + qwarp;
    int lp_b = c_first / m, lp_mm = c_first % m;                 // of the next tile load_point is asked for
    int ep_b = lp_b, ep_mm = lp_mm;                              // of the next tile epilogue3 is asked for
    // (the subtraction neighbour - centre is left to the consumer one phase later: done here, the FADD would sit on the
    // loads' scoreboard and hold up everything behind it -- 7.6 % of the stall samples of the previous version)
    auto load_point = [&](int src, float (&g)[6]) {              // tiles in visiting order, one call per tile
#pragma unroll
        for (int i = 0; i < 6; ++i) g[i] = 0.f;
        if (src >= 0) {
            const float* p = xyz + ((size_t)lp_b * n + src) * 3;
            const float* c = new_xyz + ((size_t)lp_b * m + lp_mm) * 3;
            g[0] = __ldg(p); g[1] = __ldg(p + 1); g[2] = __ldg(p + 2);
            g[3] = __ldg(c); g[4] = __ldg(c + 1); g[5] = __ldg(c + 2);
        }
        advance(lp_b, lp_mm);
    };
    const int tstep = gridDim.x;
    float ng[6];
    load_point(load_index(blockIdx.x), ng);
    int next_src = load_index(blockIdx.x + tstep);
    const int row = threadIdx.x & 127;
    constexpr int kHalf1 = kC1 / 2;           // channels of layer 1 per thread

    // layer 1 of a tile on the CUDA cores (this thread's half of the channels), result kept in registers: the A buffer
    // may still feed the tensor core
    auto layer1 = [&](float g0, float g1, float g2, float (&h)[kHalf1]) {
        const unsigned long long G0 = pack2(g0, g0), G1 = pack2(g1, g1), G2 = pack2(g2, g2);
        const float4* wp = reinterpret_cast<const float4*>(sW1) + half * kHalf1;     // two float4 per channel pair
#pragma unroll
        for (int i = 0; i < kHalf1 / 2; ++i) {
            const float4 wa = wp[2 * i], wb = wp[2 * i + 1];
            unsigned long long acc = pack2(wb.z, wb.w);                               // same order as the scalar chain:
            acc = fma2(pack2(wa.x, wa.y), G0, acc);                                   // b + w0 g0, + w1 g1, + w2 g2
            acc = fma2(pack2(wa.z, wa.w), G1, acc);
            acc = fma2(pack2(wb.x, wb.y), G2, acc);
            float a0, a1;
            unpack2(acc, a0, a1);
            h[2 * i] = fmaxf(a0, 0.f);
            h[2 * i + 1] = fmaxf(a1, 0.f);
        }
    };
    auto write_a = [&](const float (&h)[kHalf1]) {
#pragma unroll
        for (int kc = 0; kc < kHalf1 / 8; ++kc) {
            float x[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = h[kc * 8 + i];
            store_chunk(smem, row, half * (kHalf1 / 8) + kc, x);
        }
    };
    // Hand-over of a finished A operand to the tensor core WITHOUT a CTA barrier: every warp counts itself in on a shared
    // counter once its lanes' stores are fenced, and the warp that completes the round issues the MMAs.  Nobody waits for
    // anybody here -- the data hazards (A still being read, accumulators still being drained) are all covered by the
    // MMA-completion mbarrier every warp waits on before it touches A or TMEM again -- so the warps drift apart and the
    // CUDA-core phases of some overlap the tensor-core phases of others (BAR.SYNC and the instruction behind it held
    // 12 % of the stall samples of the version with __syncthreads + thread 0).
    unsigned arrive_target = 0;
    auto publish_and_issue = [&](int w_hi, int w_lo, unsigned tmem_d, unsigned idesc) {
        // generic-proxy smem writes -> visible to the tensor core (async proxy); this thread's TMEM reads are complete
        // (tcgen05.wait::ld) and ordered before the MMAs that overwrite the accumulators
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        arrive_target += kTcThreads / 32;
        if (lane == 0) {
            unsigned before;
            asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(before) : "r"(s_addr(arrive_cnt)) : "memory");
            if (before + 1 == arrive_target) {                    // 12 MMAs: K = 64 in steps of 16, hi*hi + hi*lo + lo*hi
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const unsigned long long ah = umma_desc(sbase + kOffAh + j * 256), al = umma_desc(sbase + kOffAl + j * 256);
                    const unsigned long long wh = umma_desc(sbase + w_hi + j * 256), wl = umma_desc(sbase + w_lo + j * 256);
                    umma_f16(tmem_d, ah, wh, idesc, j > 0);
                    umma_f16(tmem_d, ah, wl, idesc, 1);
                    umma_f16(tmem_d, al, wh, idesc, 1);
                }
                umma_commit(bar);
            }
        }
    };
    // epilogue of layer 3 = bias + ReLU + max over the 32 neighbours of a centre, i.e. over the 32 TMEM lanes of this warp.
    // max_j relu(v_j + b) == relu(max_j v_j + b) bit for bit (rounding is monotone), so the maximum is taken on the raw
    // accumulators.  tcgen05.ld.16x256b hands every thread FOUR rows of the same two columns per 8-column group (rows
    // t/4, t/4+8 of each 16-lane half; columns 2(t%4), 2(t%4)+1 -- the mma accumulator fragment layout, checked on the
    // device with profiles/tools/tmem_map.cu): three in-thread FMNMX per column, then a reduce-scatter butterfly over
    // the 8 threads that share t%4 (7 shuffles per 32 columns) instead of one REDUX per column and lane.
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
    const int colsel = 8 * (2 * (int)b4 + (int)b3) + 2 * (int)(lane & 3) + (int)b2;
    auto epilogue3 = [&](int centre) {                           // tiles in visiting order, one call per tile
        const bool live = centre < n_centres;
        float* const obase = out + (size_t)ep_b * kC3 * m + ep_mm;
        advance(ep_b, ep_mm);
        unsigned ra[2][16], rc[2][16];
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {                          // this warp's half of the 128 channels: four loads in flight
            const unsigned t0 = tmem_base + ((unsigned)(qwarp * 32) << 16) + kC2 + (half * 2 + cc) * 32;
            tmem_ld_16x256b_x4(t0, ra[cc]);
            tmem_ld_16x256b_x4(t0 + (16u << 16), rc[cc]);
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
            const int chunk = half * 2 + cc;
            const unsigned (&a)[16] = ra[cc];
            const unsigned (&c)[16] = rc[cc];
            float v[8];
#pragma unroll
            for (int rep = 0; rep < 4; ++rep) {
#pragma unroll
                for (int e = 0; e < 2; ++e)
                    v[2 * rep + e] = fmaxf(fmaxf(__uint_as_float(a[4 * rep + e]), __uint_as_float(a[4 * rep + 2 + e])),
                                           fmaxf(__uint_as_float(c[4 * rep + e]), __uint_as_float(c[4 * rep + 2 + e])));
            }
            float w4[4], w2[2];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float mine = b4 ? v[4 + i] : v[i], send = b4 ? v[i] : v[4 + i];
                w4[i] = fmaxf(mine, __shfl_xor_sync(0xffffffffu, send, 16));
            }
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const float mine = b3 ? w4[2 + e] : w4[e], send = b3 ? w4[e] : w4[2 + e];
                w2[e] = fmaxf(mine, __shfl_xor_sync(0xffffffffu, send, 8));
            }
            const float mine = b2 ? w2[1] : w2[0], send = b2 ? w2[0] : w2[1];
            const float top = fmaxf(mine, __shfl_xor_sync(0xffffffffu, send, 4));
            const int col = chunk * 32 + colsel;
            if (live) obase[(size_t)col * m] = fmaxf(top + sB3[col], 0.f);
        }
    };

    // ---- software pipeline: the tensor core works on one layer while the CUDA cores do the other layers' share ----
    //   issue L2(t) | epilogue 3(t-1) | wait L2 | epilogue 2(t) -> A | issue L3(t) | layer 1(t+1) in registers | wait L3 |
    //   A <- layer 1(t+1) | issue L2(t+1) | ...
    int tile = blockIdx.x;
    if (tile < n_tiles) {
        float h[kHalf1];
        layer1(__fsub_rn(ng[0], ng[3]), __fsub_rn(ng[1], ng[4]), __fsub_rn(ng[2], ng[5]), h);
        const int src1 = next_src;
        next_src = load_index(tile + 2 * tstep);
        load_point(src1, ng);
        write_a(h);
        publish_and_issue(kOffW2h, kOffW2l, tmem_base, idesc2);
    }
    int prev_centre = -1;
    for (; tile < n_tiles; tile += tstep) {
        if (prev_centre >= 0) epilogue3(prev_centre);             // under the layer-2 MMAs of this tile
        bar_wait(bar, phase);
        phase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // ---- epilogue 2: bias + ReLU, re-split, becomes the A operand of layer 3 -----------------
        {
            unsigned r0[16], r1[16];                              // this thread's 32 of the 64 channels: both loads in flight
            tmem_ld16_nowait(tmem_lane + (half * 2) * 16, r0);
            tmem_ld16_nowait(tmem_lane + (half * 2 + 1) * 16, r1);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int qq = 0; qq < 2; ++qq) {
                const int q = half * 2 + qq;
#pragma unroll
                for (int hlf = 0; hlf < 2; ++hlf) {
                    float x[8];
                    const float4* bp = reinterpret_cast<const float4*>(sB2 + q * 16 + hlf * 8);
                    const float4 ba = bp[0], bb = bp[1];
                    const float bias[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
                    for (int i = 0; i < 8; i += 2) {
                        const unsigned u0 = qq ? r1[hlf * 8 + i] : r0[hlf * 8 + i], u1 = qq ? r1[hlf * 8 + i + 1] : r0[hlf * 8 + i + 1];
                        float s0, s1;
                        unpack2(add2(pack2(__uint_as_float(u0), __uint_as_float(u1)), pack2(bias[i], bias[i + 1])), s0, s1);
                        x[i] = fmaxf(s0, 0.f);
                        x[i + 1] = fmaxf(s1, 0.f);
                    }
                    store_chunk(smem, row, q * 2 + hlf, x);
                }
            }
        }
        publish_and_issue(kOffW3h, kOffW3l, tmem_base + kC2, idesc3);
        prev_centre = tile * 4 + qwarp;
        // ---- layer 1 of the next tile, under the layer-3 MMAs ---------------------------------------
        const bool more = tile + tstep < n_tiles;
        float h[kHalf1];
        if (more) {
            layer1(__fsub_rn(ng[0], ng[3]), __fsub_rn(ng[1], ng[4]), __fsub_rn(ng[2], ng[5]), h);
            const int src1 = next_src;
            next_src = load_index(tile + 3 * tstep);
            load_point(src1, ng);
        }
        bar_wait(bar, phase);                                    // layer 3 done: the A buffer is free, D3 is complete
        phase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (more) {
            write_a(h);
            publish_and_issue(kOffW2h, kOffW2l, tmem_base, idesc2);
        }
    }
    if (prev_centre >= 0) epilogue3(prev_centre);
    // ---- teardown ----------------------------------------------------------------------------------
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

int launch_shared_mlp_tc(const float* xyz, const int* idx, const float* new_xyz, int b, int n, int m, int k,
                         const float* W1, const float* B1, const float* W2, const float* B2, const float* W3,
                         const float* B3, float* out, cudaStream_t st) {
    (void)k;
    LIDAR_CUDA_TRY(cudaFuncSetAttribute(shared_mlp_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmem));
    const int n_centres = b * m;
    const int n_tiles = (n_centres + 3) / 4;
    int grid = sm_count() * 2;   // 2 CTAs per SM: 2 x 256 TMEM columns, 2 x 82 KB of shared memory, 2 x 256 threads
    if (grid > n_tiles) grid = n_tiles;
    shared_mlp_tc_kernel<<<grid, kTcThreads, kTcSmem, st>>>(xyz, idx, new_xyz, n, m, n_centres, W1, B1, W2, B2, W3, B3, out);
    LIDAR_CHECK_LAUNCH();
    return LIDAR_OK;
}

}  // namespace lidar
