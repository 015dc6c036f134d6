// K9 / K10 — lattice flow field, bottleneck scoring, radius counts and frame-to-frame flow.
//
//   lidar_flow_field        _generate_simulated_flow      models/crowd_flow_model.py:88-184 (variant A)
//                                                         app_simplified.py:348-411          (variant B)
//   lidar_flow_bottlenecks  _identify_bottlenecks         models/crowd_flow_model.py:186-279
//   lidar_flow_box_max      box-neighbour rule            app_simplified.py:427-447
//   lidar_radius_count      KDTree.query_radius count     app_simplified.py:266-281 (r = 2 m per cell)
//   lidar_frame_flow_match / lidar_frame_flow_field       NEW op, SURVEY.md Appendix B.3
//
// The lattice is tiny (<= ~1e5 nodes) so these kernels are latency-, not bandwidth-bound; what matters
// is that the whole analyze() call stays on the device and that every inclusive / strict compare is
// taken in fp64 on the same expression the reference evaluates (lattice distances hit exactly 3.0 and
// 5.0).
#include "common.cuh"

namespace lidar {

struct FlowParams {
    double exit_x, exit_y;
    double freq, amp;          // angle_mod = sin(x*freq) * cos(y*freq) * amp
    double bx[3], by[3];       // slow-down discs (host-drawn from the legacy MT19937 stream)
    int n_discs;
};

__device__ __forceinline__ unsigned long long nonneg_f64_bits(double v) {
    return (unsigned long long)__double_as_longlong(v);   // order preserving for v >= 0
}

// pass 1: un-scaled vectors, and the maximum magnitude (atomicMax on the bit pattern)
__global__ void flow_vectors_kernel(const double* __restrict__ xg, int nx, const double* __restrict__ yg, int ny,
                                    FlowParams P, double* __restrict__ pos, double* __restrict__ vec,
                                    unsigned long long* __restrict__ max_bits) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double mag = 0.0;
    if (i < nx * ny) {
        const double x = xg[i % nx], y = yg[i / nx];   // meshgrid(x_grid, y_grid).ravel(): y outer
        double dx = __dsub_rn(P.exit_x, x), dy = __dsub_rn(P.exit_y, y);
        const double dist = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
        double vx = 0.0, vy = 0.0;
        if (dist > 0.0) {
            dx = __ddiv_rn(dx, dist);
            dy = __ddiv_rn(dy, dist);
            const double ang = __dmul_rn(__dmul_rn(sin(__dmul_rn(x, P.freq)), cos(__dmul_rn(y, P.freq))), P.amp);
            const double c = cos(ang), s = sin(ang);
            vx = __dsub_rn(__dmul_rn(dx, c), __dmul_rn(dy, s));
            vy = __dadd_rn(__dmul_rn(dx, s), __dmul_rn(dy, c));
        }
        for (int k = 0; k < P.n_discs; ++k) {
            const double ex = __dsub_rn(x, P.bx[k]), ey = __dsub_rn(y, P.by[k]);
            const double d = sqrt(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)));
            if (d < 3.0) {
                const double f = __ddiv_rn(d, 3.0);
                vx = __dmul_rn(vx, f);
                vy = __dmul_rn(vy, f);
            }
        }
        pos[2 * i] = x; pos[2 * i + 1] = y;
        vec[2 * i] = vx; vec[2 * i + 1] = vy;
        mag = sqrt(__dadd_rn(__dmul_rn(vx, vx), __dmul_rn(vy, vy)));
    }
    unsigned long long b = nonneg_f64_bits(mag);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long t = __shfl_xor_sync(0xffffffffu, b, o);
        b = t > b ? t : b;
    }
    if (lane_id() == 0 && b) atomicMax(max_bits, b);
}

// pass 2: scale, magnitudes (optionally clipped), and the three sums behind avg_speed / direction
__global__ void flow_scale_kernel(int n, double span, int clip, double lo, double hi,
                                  const unsigned long long* __restrict__ max_bits, double* __restrict__ vec,
                                  double* __restrict__ mag_out, double* __restrict__ sums3) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const double top = __longlong_as_double((long long)*max_bits);
    const double scale = top > 0.0 ? __ddiv_rn(span, top) : 1.0;
    double m = 0.0, vx = 0.0, vy = 0.0;
    if (i < n) {
        vx = __dmul_rn(vec[2 * i], scale);
        vy = __dmul_rn(vec[2 * i + 1], scale);
        vec[2 * i] = vx; vec[2 * i + 1] = vy;
        m = sqrt(__dadd_rn(__dmul_rn(vx, vx), __dmul_rn(vy, vy)));
        if (clip) m = m < lo ? lo : (m > hi ? hi : m);
        mag_out[i] = m;
    }
    m = warp_sum(m); vx = warp_sum(vx); vy = warp_sum(vy);
    if (lane_id() == 0) {
        atomicAdd(&sums3[0], m);     // fp64 atomics: only feeds avg_speed / the compass octant (rtol 1e-3)
        atomicAdd(&sums3[1], vx);
        atomicAdd(&sums3[2], vy);
    }
}

// variant A bottleneck severity per lattice node (0 when the node does not qualify)
__global__ void flow_bottleneck_kernel(int nx, int ny, const double* __restrict__ pos, const double* __restrict__ vec,
                                       const double* __restrict__ mag, double* __restrict__ severity) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nx * ny) return;
    double sev = 0.0;
    const double m = mag[i];
    if (m <= 0.5) {
        const int ci = i % nx, cj = i / nx;
        const double px = pos[2 * i], py = pos[2 * i + 1];
        int n_near = 0, n_far = 0;
        double s_near = 0.0, s_far = 0.0, conv = 0.0;
        const int R = 7;   // lattice pitch is ~1 m: radius 5 m is at most 6 index steps (+1 of slack)
        for (int dj = -R; dj <= R; ++dj) {
            const int j = cj + dj;
            if (j < 0 || j >= ny) continue;
            for (int di = -R; di <= R; ++di) {
                const int k = ci + di;
                if (k < 0 || k >= nx) continue;
                const int q = j * nx + k;
                const double dx = __dsub_rn(px, pos[2 * q]), dy = __dsub_rn(py, pos[2 * q + 1]);
                const double r2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
                if (r2 <= 9.0) {
                    ++n_near;
                    s_near += mag[q];
                    const double nrm = sqrt(r2);
                    if (nrm > 0.0) {
                        const double ux = __ddiv_rn(dx, nrm), uy = __ddiv_rn(dy, nrm);
                        const double dot = __dadd_rn(__dmul_rn(ux, vec[2 * q]), __dmul_rn(uy, vec[2 * q + 1]));
                        if (dot > 0.0) conv += dot;
                    }
                } else if (r2 <= 25.0) {
                    ++n_far;
                    s_far += mag[q];
                }
            }
        }
        if (n_near >= 5 && n_far >= 3) {
            const double grad = __dsub_rn(__ddiv_rn(s_far, (double)n_far), __ddiv_rn(s_near, (double)n_near));
            const double c = __ddiv_rn(conv, (double)n_near);
            sev = __ddiv_rn(__dadd_rn(__dmul_rn(grad, 5.0), __dmul_rn(c, 5.0)), 2.0);
        }
    }
    severity[i] = sev;
}

// variant B: max speed inside the open +-3 m box around every slow node (-1 elsewhere)
__global__ void flow_box_max_kernel(int nx, int ny, const double* __restrict__ pos, const double* __restrict__ mag,
                                    double slow_below, double* __restrict__ box_max) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nx * ny) return;
    double best = -1.0;
    if (mag[i] < slow_below) {
        const int ci = i % nx, cj = i / nx;
        const double px = pos[2 * i], py = pos[2 * i + 1];
        const int R = 4;
        for (int dj = -R; dj <= R; ++dj) {
            const int j = cj + dj;
            if (j < 0 || j >= ny) continue;
            for (int di = -R; di <= R; ++di) {
                const int k = ci + di;
                if (k < 0 || k >= nx) continue;
                const int q = j * nx + k;
                if (fabs(__dsub_rn(pos[2 * q], px)) < 3.0 && fabs(__dsub_rn(pos[2 * q + 1], py)) < 3.0)
                    best = fmax(best, mag[q]);
            }
        }
    }
    box_max[i] = best;
}

// #centres with rdist <= r2 per query (KDTree.query_radius(..., count_only) semantics, inclusive)
__global__ void radius_count_kernel(const double* __restrict__ centres, int n_centres, const double* __restrict__ qx,
                                    int nqx, const double* __restrict__ qy, int nqy, double r2,
                                    int* __restrict__ counts /*[nqy][nqx]*/) {
    extern __shared__ double s_c[];
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = q < nqx * nqy;
    const double x = live ? qx[q % nqx] : 0.0, y = live ? qy[q / nqx] : 0.0;
    int cnt = 0;
    for (int base = 0; base < n_centres; base += 1024) {
        const int chunk = n_centres - base < 1024 ? n_centres - base : 1024;
        __syncthreads();
        for (int t = threadIdx.x; t < 2 * chunk; t += blockDim.x) s_c[t] = centres[2 * base + t];
        __syncthreads();
        if (live)
            for (int c = 0; c < chunk; ++c) {
                const double dx = __dsub_rn(x, s_c[2 * c]), dy = __dsub_rn(y, s_c[2 * c + 1]);
                cnt += __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)) <= r2;
            }
    }
    if (live) counts[q] = cnt;
}

// B.3: nearest previous centroid (fp32, lowest index on ties), gated.  One warp per current centroid: the lanes
// stride over the previous frame, then a lexicographic (distance, index) minimum across the warp -- the same
// winner as the index-order walk (a single thread per centroid walking ~1500 dependent loads took 69 us).
__global__ void __launch_bounds__(256)
frame_flow_match_kernel(const float* __restrict__ prev, int n_prev, const float* __restrict__ cur,
                        int n_cur, float dt, float gate2, int* __restrict__ match,
                        float* __restrict__ vel) {
    const int j = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (j >= n_cur) return;                            // whole warps leave together
    const float cx = cur[2 * j], cy = cur[2 * j + 1];
    float best = INFINITY;
    int bi = 0x7fffffff;
    for (int i = (int)lane_id(); i < n_prev; i += 32) {
        const float2 p = reinterpret_cast<const float2*>(prev)[i];
        const float dx = __fsub_rn(cx, p.x), dy = __fsub_rn(cy, p.y);
        const float d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
        if (d2 < best) { best = d2; bi = i; }          // ascending i per lane: the first minimum is kept
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (lane_id() != 0) return;
    float vx = 0.f, vy = 0.f;
    if (bi != 0x7fffffff && best <= gate2) {
        vx = __fdiv_rn(__fsub_rn(cx, prev[2 * bi]), dt);
        vy = __fdiv_rn(__fsub_rn(cy, prev[2 * bi + 1]), dt);
    } else {
        bi = -1;
    }
    match[j] = bi;
    vel[2 * j] = vx; vel[2 * j + 1] = vy;
}

// Mean velocity of the matched people within `radius` of every lattice node, people taken in index order (the sum
// order of the frozen restatement).  The people are staged in shared memory tile by tile, so the per-node loop is
// arithmetic only (one thread walking ~1500 x 5 dependent global loads took 219 us for a 12 k-node lattice).
constexpr int kFieldThreads = 64;
constexpr int kFieldTile = 1024;
__global__ void __launch_bounds__(kFieldThreads)
frame_flow_field_kernel(const double* __restrict__ lattice, int g, const float* __restrict__ cur,
                        const int* __restrict__ match, const float* __restrict__ vel, int n_cur,
                        double r2, double* __restrict__ vec, double* __restrict__ mag) {
    __shared__ float4 s_p[kFieldTile];                 // {x, y, vx, vy}; unmatched people get x = NaN (never inside)
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < g;
    const double x = live ? lattice[2 * i] : 0.0, y = live ? lattice[2 * i + 1] : 0.0;
    double sx = 0.0, sy = 0.0;
    int cnt = 0;
    for (int base = 0; base < n_cur; base += kFieldTile) {
        const int chunk = n_cur - base < kFieldTile ? n_cur - base : kFieldTile;
        __syncthreads();
        for (int t = threadIdx.x; t < chunk; t += kFieldThreads) {
            const int j = base + t;
            const bool ok = match[j] >= 0;
            s_p[t] = make_float4(ok ? cur[2 * j] : __int_as_float(0x7fc00000), cur[2 * j + 1], vel[2 * j], vel[2 * j + 1]);
        }
        __syncthreads();
        if (live) {
#pragma unroll 4
            for (int t = 0; t < chunk; ++t) {
                const float4 p = s_p[t];
                const double dx = __dsub_rn(x, (double)p.x), dy = __dsub_rn(y, (double)p.y);
                if (__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)) <= r2) {
                    ++cnt;
                    sx += (double)p.z;
                    sy += (double)p.w;
                }
            }
        }
    }
    if (!live) return;
    const double vx = cnt ? __ddiv_rn(sx, (double)cnt) : 0.0, vy = cnt ? __ddiv_rn(sy, (double)cnt) : 0.0;
    vec[2 * i] = vx; vec[2 * i + 1] = vy;
    mag[i] = sqrt(__dadd_rn(__dmul_rn(vx, vx), __dmul_rn(vy, vy)));
}


// np.vstack([X.ravel(), Y.ravel()]).T of X, Y = np.meshgrid(x_grid, y_grid) (models/crowd_flow_model.py:110-111): node
// i = iy * nx + ix is (x_grid[ix], y_grid[iy]).  Copies, no arithmetic.
__global__ void lattice_kernel(const double* __restrict__ xg, int nx, const double* __restrict__ yg, int ny,
                               double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nx * ny) return;
    reinterpret_cast<double2*>(out)[i] = make_double2(xg[i % nx], yg[i / nx]);
}

// plot_crowd_metrics join (utils/visualization.py:306-326): cKDTree(density cell centres).query(flow nodes, k=1) and the
// congestion-risk columns built from it.  The cell centres are the rectilinear grid repeat(grid_x, ny) x tile(grid_y, nx)
// of CrowdDensityModel.analyze (models/crowd_density_model.py:57-59), so the nearest centre is found per axis: the two
// grid lines around the node (binary search), then the four candidate cells are compared on the fp64 squared distance
// dx*dx + dy*dy the KD-tree itself minimises.  An exact tie goes to the lowest flat index (cKDTree breaks ties by
// traversal order; the DISTANCE is unique either way).
__device__ __forceinline__ int lower_line(const double* __restrict__ g, int n, double x) {
    int lo = 0, hi = n;                 // first index with g[idx] > x, minus one, clamped to [0, n-1]
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (g[mid] <= x) lo = mid + 1; else hi = mid;
    }
    const int k = lo - 1;
    return k < 0 ? 0 : k;
}

__global__ void __launch_bounds__(128)
nearest_cell_kernel(const double* __restrict__ nodes_xy, int n_nodes, const double* __restrict__ gx, int nx,
                    const double* __restrict__ gy, int ny, const double* __restrict__ density_flat,
                    const double* __restrict__ speed, long long* __restrict__ index, double* __restrict__ distance,
                    double* __restrict__ density_at, double* __restrict__ risk, unsigned long long* __restrict__ risk_max_bits) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double r = 0.0;
    if (i < n_nodes) {
        const double x = nodes_xy[2 * i], y = nodes_xy[2 * i + 1];
        const int kx = lower_line(gx, nx, x), ky = lower_line(gy, ny, y);
        double best = INFINITY;
        long long arg = 0;
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            const int cx = kx + a < nx ? kx + a : nx - 1;
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                const int cy = ky + b < ny ? ky + b : ny - 1;
                const double dx = __dsub_rn(x, gx[cx]), dy = __dsub_rn(y, gy[cy]);
                const double d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
                const long long flat = (long long)cx * ny + cy;
                if (d2 < best || (d2 == best && flat < arg)) { best = d2; arg = flat; }
            }
        }
        index[i] = arg;
        distance[i] = __dsqrt_rn(best);
        if (density_flat) {
            const double d = density_flat[arg];
            density_at[i] = d;
            r = __ddiv_rn(d, __dadd_rn(speed[i], 0.1));           // congestion_risk = density / (speed + 0.1)
            risk[i] = r;
        }
    }
    if (risk_max_bits) {
        // max over non-negative doubles = max over their bit patterns (NaN aside: density and speed are finite)
        r = r > 0.0 ? r : 0.0;
        r = warp_max(r);
        if (lane_id() == 0 && r > 0.0) atomicMax(risk_max_bits, (unsigned long long)__double_as_longlong(r));
    }
}

}  // namespace lidar

using namespace lidar;

extern "C" {

int lidar_flow_field(const double* d_xgrid, int nx, const double* d_ygrid, int ny, double exit_x, double exit_y,
                     double freq, double amp, const double* h_discs_xy, int n_discs, double speed_span, int clip,
                     double clip_lo, double clip_hi, double* d_positions, double* d_vectors, double* d_magnitudes,
                     double* d_sums4, void* stream) {
    LIDAR_REQUIRE(nx > 0 && ny > 0 && (int64_t)nx * ny < (1ll << 30), LIDAR_ERR_INVALID, "lidar_flow_field: bad lattice");
    LIDAR_REQUIRE(n_discs >= 0 && n_discs <= 3 && (n_discs == 0 || h_discs_xy), LIDAR_ERR_INVALID,
                  "lidar_flow_field: at most 3 discs");
    LIDAR_REQUIRE(d_xgrid && d_ygrid && d_positions && d_vectors && d_magnitudes && d_sums4, LIDAR_ERR_INVALID,
                  "lidar_flow_field: NULL argument");
    FlowParams P;
    P.exit_x = exit_x; P.exit_y = exit_y; P.freq = freq; P.amp = amp; P.n_discs = n_discs;
    for (int k = 0; k < 3; ++k) { P.bx[k] = k < n_discs ? h_discs_xy[2 * k] : 0.0; P.by[k] = k < n_discs ? h_discs_xy[2 * k + 1] : 0.0; }
    cudaStream_t st = as_stream(stream);
    const int n = nx * ny;
    LIDAR_CUDA_TRY(cudaMemsetAsync(d_sums4, 0, sizeof(double) * 4, st));
    unsigned long long* max_bits = reinterpret_cast<unsigned long long*>(d_sums4 + 3);
    flow_vectors_kernel<<<(n + 127) / 128, 128, 0, st>>>(d_xgrid, nx, d_ygrid, ny, P, d_positions, d_vectors, max_bits);
    LIDAR_CHECK_LAUNCH();
    flow_scale_kernel<<<(n + 127) / 128, 128, 0, st>>>(n, speed_span, clip, clip_lo, clip_hi, max_bits, d_vectors,
                                                       d_magnitudes, d_sums4);
    LIDAR_CHECK_LAUNCH();
    return LIDAR_OK;
}

int lidar_flow_bottlenecks(int nx, int ny, const double* d_positions, const double* d_vectors,
                           const double* d_magnitudes, double* d_severity, void* stream) {
    LIDAR_REQUIRE(nx > 0 && ny > 0 && d_positions && d_vectors && d_magnitudes && d_severity, LIDAR_ERR_INVALID,
                  "lidar_flow_bottlenecks: bad argument");
    const int n = nx * ny;
    // 64-thread CTAs: a 100 m x 100 m lattice is ~10 k nodes, i.e. 160 CTAs for 148 SMs instead of 80
    flow_bottleneck_kernel<<<(n + 63) / 64, 64, 0, as_stream(stream)>>>(nx, ny, d_positions, d_vectors,
                                                                        d_magnitudes, d_severity);
    LIDAR_CHECK_LAUNCH();
    return LIDAR_OK;
}

int lidar_flow_box_max(int nx, int ny, const double* d_positions, const double* d_magnitudes, double slow_below,
                       double* d_box_max, void* stream) {
    LIDAR_REQUIRE(nx > 0 && ny > 0 && d_positions && d_magnitudes && d_box_max, LIDAR_ERR_INVALID,
                  "lidar_flow_box_max: bad argument");
    const int n = nx * ny;
    flow_box_max_kernel<<<(n + 127) / 128, 128, 0, as_stream(stream)>>>(nx, ny, d_positions, d_magnitudes, slow_below,
                                                                       d_box_max);
    LIDAR_CHECK_LAUNCH();
    return LIDAR_OK;
}

int lidar_radius_count(const double* d_centres_xy, int n_centres, const double* d_qx, int nqx, const double* d_qy,
                       int nqy, double radius, int32_t* d_counts, void* stream) {
    LIDAR_REQUIRE(n_centres >= 0 && nqx > 0 && nqy > 0 && d_qx && d_qy && d_counts, LIDAR_ERR_INVALID,
                  "lidar_radius_count: bad argument");
    LIDAR_REQUIRE(n_centres == 0 || d_centres_xy, LIDAR_ERR_INVALID, "lidar_radius_count: NULL centres");
    const int n = nqx * nqy;
    radius_count_kernel<<<(n + 127) / 128, 128, 2048 * sizeof(double), as_stream(stream)>>>(
        d_centres_xy, n_centres, d_qx, nqx, d_qy, nqy, radius * radius, d_counts);
    LIDAR_CHECK_LAUNCH();
    return LIDAR_OK;
}

int lidar_nearest_grid_cell(const double* d_nodes_xy, int n_nodes, const double* d_gx, int nx, const double* d_gy, int ny,
                            const double* d_density_flat, const double* d_speed, int64_t* d_index, double* d_distance,
                            double* d_density_at, double* d_risk, double* d_risk_max, void* stream) {
    LIDAR_REQUIRE(n_nodes >= 0 && nx > 0 && ny > 0 && d_gx && d_gy && d_index && d_distance, LIDAR_ERR_INVALID,
                  "lidar_nearest_grid_cell: bad argument");
    LIDAR_REQUIRE(!d_density_flat || (d_speed && d_density_at && d_risk && d_risk_max), LIDAR_ERR_INVALID,
                  "lidar_nearest_grid_cell: the risk columns need speed, density_at, risk and risk_max");
    if (n_nodes == 0) return LIDAR_OK;
    LIDAR_REQUIRE(d_nodes_xy != nullptr, LIDAR_ERR_INVALID, "lidar_nearest_grid_cell: NULL nodes");
    cudaStream_t st = as_stream(stream);
    if (d_risk_max) LIDAR_CUDA_TRY(cudaMemsetAsync(d_risk_max, 0, sizeof(double), st));
    nearest_cell_kernel<<<(n_nodes + 127) / 128, 128, 0, st>>>(d_nodes_xy, n_nodes, d_gx, nx, d_gy, ny, d_density_flat, d_speed,
                                                               reinterpret_cast<long long*>(d_index), d_distance, d_density_at,
                                                               d_risk, reinterpret_cast<unsigned long long*>(d_risk_max));
    LIDAR_CHECK_LAUNCH();
    return LIDAR_OK;
}

int lidar_frame_flow_match(const float* d_prev_xy, int n_prev, const float* d_cur_xy, int n_cur, float dt, float gate,
                           int32_t* d_match, float* d_velocity, void* stream) {
    LIDAR_REQUIRE(n_prev >= 0 && n_cur >= 0 && dt > 0.f && gate >= 0.f, LIDAR_ERR_INVALID, "lidar_frame_flow_match: bad argument");
    if (n_cur == 0) return LIDAR_OK;
    LIDAR_REQUIRE(d_cur_xy && d_match && d_velocity && (n_prev == 0 || d_prev_xy), LIDAR_ERR_INVALID,
                  "lidar_frame_flow_match: NULL argument");
    frame_flow_match_kernel<<<(n_cur + 7) / 8, 256, 0, as_stream(stream)>>>(d_prev_xy, n_prev, d_cur_xy, n_cur, dt,
                                                                           gate * gate, d_match, d_velocity);   // one warp per centroid
    LIDAR_CHECK_LAUNCH();
    return LIDAR_OK;
}

int lidar_frame_flow_field(const double* d_lattice_xy, int n_lattice, const float* d_cur_xy, const int32_t* d_match,
                           const float* d_velocity, int n_cur, double radius, double* d_vectors, double* d_magnitudes,
                           void* stream) {
    LIDAR_REQUIRE(n_lattice >= 0 && n_cur >= 0, LIDAR_ERR_INVALID, "lidar_frame_flow_field: bad sizes");
    if (n_lattice == 0) return LIDAR_OK;
    LIDAR_REQUIRE(d_lattice_xy && d_vectors && d_magnitudes && (n_cur == 0 || (d_cur_xy && d_match && d_velocity)),
                  LIDAR_ERR_INVALID, "lidar_frame_flow_field: NULL argument");
    frame_flow_field_kernel<<<(n_lattice + kFieldThreads - 1) / kFieldThreads, kFieldThreads, 0, as_stream(stream)>>>(
        d_lattice_xy, n_lattice, d_cur_xy, d_match, d_velocity, n_cur, radius * radius, d_vectors, d_magnitudes);
    LIDAR_CHECK_LAUNCH();
    return LIDAR_OK;
}

int lidar_frame_flow(const float* d_prev_xy, int n_prev, const float* d_cur_xy, int n_cur, float dt, float gate,
                     const double* d_x_grid, int nx, const double* d_y_grid, int ny, double radius, int32_t* d_match,
                     float* d_velocity, double* d_lattice_xy, double* d_vectors, double* d_magnitudes, void* stream) {
    LIDAR_REQUIRE(nx >= 0 && ny >= 0 && (int64_t)nx * ny < (1ll << 31), LIDAR_ERR_INVALID, "lidar_frame_flow: bad lattice");
    const int g = nx * ny;
    cudaStream_t st = as_stream(stream);
    if (g > 0) {
        LIDAR_REQUIRE(d_x_grid && d_y_grid && d_lattice_xy, LIDAR_ERR_INVALID, "lidar_frame_flow: NULL lattice argument");
        lattice_kernel<<<(g + 255) / 256, 256, 0, st>>>(d_x_grid, nx, d_y_grid, ny, d_lattice_xy);
        LIDAR_CHECK_LAUNCH();
    }
    if (n_cur > 0 && d_velocity) LIDAR_CUDA_TRY(cudaMemsetAsync(d_velocity, 0, sizeof(float) * 2 * (size_t)n_cur, st));
    int rc = lidar_frame_flow_match(d_prev_xy, n_prev, d_cur_xy, n_cur, dt, gate, d_match, d_velocity, stream);
    if (rc != LIDAR_OK) return rc;
    return lidar_frame_flow_field(d_lattice_xy, g, d_cur_xy, d_match, d_velocity, n_cur, radius, d_vectors, d_magnitudes, stream);
}

}  // extern "C"
