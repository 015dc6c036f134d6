// K7 — DBSCAN with scikit-learn-identical labels, and K8 — per-cluster centroids.
//
// Replaces sklearn.cluster.DBSCAN(eps, min_samples=5).fit(X).labels_ at
//   utils/data_processing.py:197  (variant A: on StandardScaler output, eps ~ 0.5 sigma)
//   app_simplified.py:107         (variant B: raw metres, eps = 0.3)
// and the per-cluster mean of extract_people_positions (utils/data_processing.py:251-280,
// app_simplified.py:249-255, 328-331).
//
// sklearn's labels are a pure function of the data, restated here as:
//   neighbour  <=>  rdist = ((dx*dx) + dy*dy) + dz*dz  <=  eps*eps   (fp64, inclusive, self counts)
//   core       <=>  #neighbours >= min_samples
//   clusters    =   connected components of the core-core neighbour graph, numbered by the rank of
//                   their smallest core index (dbscan_inner seeds clusters in ascending index order)
//   border      =   non-core point with a core neighbour: smallest cluster id among them
//   noise       =   -1
// GPU plan: uniform cell list with cell edge >= eps (counting sort by cell), thread per point walking
// the 3x3 (x,y) columns whose z-runs are contiguous in the sorted order; lock-free union-find that
// always links the larger root under the smaller one, so every component's root IS its smallest
// core index; an exclusive scan over "is root" in original index order numbers the clusters.
// Every output is order independent, hence deterministic although the cell sort uses atomics.
//
// Two grids, same labels:
//   general   cell edge >= eps, 3x3x3 neighbourhood, every candidate pair is tested
//   dense     cell edge = eps/sqrt(3) * (1 - 1e-6), 5x5x5 neighbourhood ("grid-exact" DBSCAN): the cell diagonal is
//             shorter than eps, so all points of a cell are mutual neighbours.  A cell with >= min_samples points
//             is all core without a single distance test, the core points of a cell always share a cluster, and
//             two cells need ONE core pair within eps to be merged -- a point skips every neighbour cell that
//             is already in its set.  A 128-beam scan has ~1000 points inside a 0.3 m ball near the sensor:
//             14.2 ms -> see DESIGN.md for a 1 M-point frame.  The margin 1e-6 dwarfs the fp64 rounding of the cell
//             assignment and of rdist (~1e-15), so "same cell => rdist <= eps^2" holds in the reference's arithmetic.
//             Certificate (tol > 0): a decision is taken on a pair that is CERTAINLY within eps whenever one
//             exists; decisions that rest on a pair inside the band, and near misses, are counted in *d_guard.
//             Pairs that are never examined (same cell, already merged cells) cannot change a certain decision.
//
// Knife-edge certificate (variant A only): X went through a scaler whose mean/scale come from a
// parallel reduction, so rdist can differ from sklearn's in the last bits.  Pairs with
// |rdist - eps^2| <= tol that could change a decision are counted in *d_guard; tests assert 0.
#include "device_scan.cuh"

namespace lidar {

constexpr int kDbThreads = 128;

struct CellGrid {
    double min[3];
    double cell;
    int g[3];
    int ncell;
    int reach;   // neighbour cells per side: 1 (general grid) or 2 (dense grid)
    int dense;
};

__device__ __forceinline__ int cell_coord(double p, double mn, double cell, int g) {
    int c = (int)floor(__ddiv_rn(__dsub_rn(p, mn), cell));
    return c < 0 ? 0 : (c >= g ? g - 1 : c);
}
__device__ __forceinline__ int cell_id(const CellGrid& G, int cx, int cy, int cz) {
    return (cx * G.g[1] + cy) * G.g[2] + cz;
}

__global__ void db_cell_assign(const double* __restrict__ pts, int m, CellGrid G, int* __restrict__ cell,
                               int* __restrict__ slot, unsigned* __restrict__ cell_count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int cx = cell_coord(pts[3 * i], G.min[0], G.cell, G.g[0]);
    const int cy = cell_coord(pts[3 * i + 1], G.min[1], G.cell, G.g[1]);
    const int cz = cell_coord(pts[3 * i + 2], G.min[2], G.cell, G.g[2]);
    const int c = cell_id(G, cx, cy, cz);
    cell[i] = c;
    slot[i] = (int)atomicAdd(&cell_count[c], 1u);
}

__global__ void db_scatter(const double* __restrict__ pts, int m, const int* __restrict__ cell,
                           const int* __restrict__ slot, const unsigned* __restrict__ cell_start,
                           int* __restrict__ sidx, int* __restrict__ scell, double* __restrict__ sx,
                           double* __restrict__ sy, double* __restrict__ sz, int* __restrict__ parent) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int c = cell[i];
    const int pos = (int)cell_start[c] + slot[i];
    sidx[pos] = i;
    scell[pos] = c;
    sx[pos] = pts[3 * i];
    sy[pos] = pts[3 * i + 1];
    sz[pos] = pts[3 * i + 2];
    parent[i] = i;
}

// Walk every candidate j of sorted point `pos` (the (2R+1)^3 cell neighbourhood); f(j, rdist) returns
// false to stop early.
template <class F>
__device__ __forceinline__ void for_each_candidate(const CellGrid& G, const unsigned* __restrict__ cell_start,
                                                   const double* __restrict__ sx, const double* __restrict__ sy,
                                                   const double* __restrict__ sz, int c, double x, double y,
                                                   double z, F&& f) {
    const int cz = c % G.g[2];
    const int t = c / G.g[2];
    const int cy = t % G.g[1];
    const int cx = t / G.g[1];
    const int R = G.reach;
    const int z0 = cz - R > 0 ? cz - R : 0;
    const int z1 = cz + R < G.g[2] - 1 ? cz + R : G.g[2] - 1;
    for (int ax = (cx - R > 0 ? cx - R : 0); ax <= (cx + R < G.g[0] - 1 ? cx + R : G.g[0] - 1); ++ax)
        for (int ay = (cy - R > 0 ? cy - R : 0); ay <= (cy + R < G.g[1] - 1 ? cy + R : G.g[1] - 1); ++ay) {
            const int j0 = (int)cell_start[cell_id(G, ax, ay, z0)];
            const int j1 = (int)cell_start[cell_id(G, ax, ay, z1) + 1];
            for (int j = j0; j < j1; ++j) {
                const double dx = __dsub_rn(x, sx[j]), dy = __dsub_rn(y, sy[j]), dz = __dsub_rn(z, sz[j]);
                double r = __dmul_rn(dx, dx);
                r = __dadd_rn(r, __dmul_rn(dy, dy));
                r = __dadd_rn(r, __dmul_rn(dz, dz));
                if (!f(j, r)) return;
            }
        }
}

// squared distance from p to the box of cell (ax, ay, az), shrunk by a relative 1e-9 so that rounding in the
// cell assignment can only make the bound smaller (a pruned cell really has no point within eps)
__device__ __forceinline__ double cell_box_dist2(const CellGrid& G, int ax, int ay, int az, double x, double y, double z) {
    const double lo_x = G.min[0] + ax * G.cell, lo_y = G.min[1] + ay * G.cell, lo_z = G.min[2] + az * G.cell;
    const double dx = fmax(0.0, fmax(lo_x - x, x - (lo_x + G.cell)));
    const double dy = fmax(0.0, fmax(lo_y - y, y - (lo_y + G.cell)));
    const double dz = fmax(0.0, fmax(lo_z - z, z - (lo_z + G.cell)));
    return (dx * dx + dy * dy + dz * dz) * (1.0 - 1e-9) - 1e-300;
}
__device__ __forceinline__ double rdist_of(double x, double y, double z, double qx, double qy, double qz) {
    const double dx = __dsub_rn(x, qx), dy = __dsub_rn(y, qy), dz = __dsub_rn(z, qz);
    double r = __dmul_rn(dx, dx);
    r = __dadd_rn(r, __dmul_rn(dy, dy));
    return __dadd_rn(r, __dmul_rn(dz, dz));
}
// first core point of the sorted run [j0, j1), -1 if none (a full cell is all core: one load)
__device__ __forceinline__ int first_core(const uint8_t* __restrict__ core_s, int j0, int j1) {
    for (int j = j0; j < j1; ++j)
        if (core_s[j]) return j;
    return -1;
}

__global__ void __launch_bounds__(kDbThreads)
db_core(int m, CellGrid G, const unsigned* __restrict__ cell_start, const int* __restrict__ scell,
        const int* __restrict__ sidx, const double* __restrict__ sx, const double* __restrict__ sy,
        const double* __restrict__ sz, double eps2, double tol, int min_samples, uint8_t* __restrict__ core_s,
        uint8_t* __restrict__ core_o, unsigned long long* __restrict__ guard, int* __restrict__ crep) {
    const int pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= m) return;
    if (G.dense) {
        // every point of the cell is a neighbour (cell diagonal < eps): a full cell is all core
        const int c = scell[pos];
        const int own = (int)(cell_start[c + 1] - cell_start[c]);
        bool core = own >= min_samples;
        if (!core) {
            // the own cell counts without a test; then the other cells, nearest columns first, pruned by their
            // box distance, until min_samples neighbours are certain (tol == 0 on this grid: no band)
            const double x = sx[pos], y = sy[pos], z = sz[pos];
            const int cz = c % G.g[2];
            const int t = c / G.g[2];
            const int cy = t % G.g[1];
            const int cx = t / G.g[1];
            const int z0 = cz - 2 > 0 ? cz - 2 : 0;
            const int z1 = cz + 2 < G.g[2] - 1 ? cz + 2 : G.g[2] - 1;
            // same-cell pairs are farther than 1e-6*eps^2 from the threshold: never inside the tol band
            int cnt = own, band_in = 0, band_out = 0;
            bool certain = false;
            // 25 column offsets in rings of growing Chebyshev / Manhattan distance, packed as (dx+2)*5 + (dy+2)
            constexpr unsigned char order[25] = {12, 7, 11, 13, 17, 6, 8, 16, 18, 2, 10, 14, 22, 1, 3, 5, 9, 15, 19, 21, 23, 0, 4, 20, 24};
            for (int k = 0; k < 25 && !certain; ++k) {
                const int ax = cx + order[k] / 5 - 2, ay = cy + order[k] % 5 - 2;
                if (ax < 0 || ay < 0 || ax >= G.g[0] || ay >= G.g[1]) continue;
                const int col = cell_id(G, ax, ay, 0);
                unsigned b1 = cell_start[col + z0];
                for (int az = z0; az <= z1 && !certain; ++az) {
                    const unsigned b0 = b1;
                    b1 = cell_start[col + az + 1];
                    if (b0 == b1 || col + az == c) continue;
                    if (cell_box_dist2(G, ax, ay, az, x, y, z) > eps2 + tol) continue;
                    for (int j0 = (int)b0; j0 < (int)b1 && !certain; j0 += 4) {   // look-ahead of four, as in db_union_dense
                        double qx[4], qy[4], qz[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const int j = j0 + k < (int)b1 ? j0 + k : (int)b1 - 1;
                            qx[k] = sx[j]; qy[k] = sy[j]; qz[k] = sz[j];
                        }
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            if (j0 + k >= (int)b1 || certain) continue;
                            const double r = rdist_of(x, y, z, qx[k], qy[k], qz[k]);
                            const bool in = r <= eps2;
                            cnt += in;
                            if (tol > 0.0 && fabs(r - eps2) <= tol) { band_in += in; band_out += !in; }
                            if (cnt - band_in >= min_samples) certain = true;   // core whatever the band pairs do
                        }
                    }
                }
            }
            core = cnt >= min_samples;
            if (((cnt - band_in) >= min_samples) != ((cnt + band_out) >= min_samples)) atomicAdd(guard, 1ull);
        }
        core_s[pos] = core;
        const int oi = sidx[pos];
        core_o[oi] = core;
        // representative of the cell = its core point with the lowest original index (kNoRep: no core point).
        // The core points of a dense cell are mutual neighbours, so one of them stands for all in the set tests.
        if (core) atomicMin(&crep[c], oi);
        return;
    }
    int cnt = 0, band_in = 0, band_out = 0;
    for_each_candidate(G, cell_start, sx, sy, sz, scell[pos], sx[pos], sy[pos], sz[pos], [&](int, double r) {
        const bool in = r <= eps2;
        cnt += in;
        if (fabs(r - eps2) <= tol && tol > 0.0) { band_in += in; band_out += !in; }
        return (cnt - band_in) < min_samples;  // certain core: stop
    });
    const bool core = cnt >= min_samples;
    if (((cnt - band_in) >= min_samples) != ((cnt + band_out) >= min_samples)) atomicAdd(guard, 1ull);
    core_s[pos] = core;
    core_o[sidx[pos]] = core;
}

constexpr int kNoRep = 0x7f7f7f7f;   // crep[] after cudaMemset(0x7f): the cell has no core point

__device__ __forceinline__ int uf_find(int* parent, int x) {
    int p = ((volatile int*)parent)[x];
    while (p != x) {
        const int gp = ((volatile int*)parent)[p];
        if (gp != p) parent[x] = gp;  // path halving (benign race: only ever moves toward the root)
        x = p;
        p = gp;
    }
    return x;
}
// find() for a set-membership QUESTION (not for linking): the walk runs on ordinary cached loads.  A stale L1 copy of a
// parent pointer is still an ancestor of the node -- links are only ever replaced by links closer to the root, roots are
// only ever hung under other roots -- so a stale walk ends at an ancestor, and the coherent find() that finishes the job
// starts from there.
__device__ __forceinline__ int ld_cached(const int* p) {
    int v;
    asm volatile("ld.global.ca.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ int uf_walk_cached(const int* parent, int x) {
    int p = ld_cached(parent + x);
    while (p != x) {
        x = p;
        p = ld_cached(parent + x);
    }
    return x;   // an ancestor of the start node: the root as far as this SM's L1 knows
}
__device__ __forceinline__ void uf_union(int* parent, int a, int b) {
    while (true) {
        a = uf_find(parent, a);
        b = uf_find(parent, b);
        if (a == b) return;
        if (a < b) { const int t = a; a = b; b = t; }  // a > b: hang a under the smaller root b
        const int old = atomicMin(&parent[a], b);
        if (old == a) return;   // a was still a root: linked
        a = old;                // somebody re-parented a meanwhile: unite that root with b
    }
}

__global__ void __launch_bounds__(kDbThreads)
db_union(int m, CellGrid G, const unsigned* __restrict__ cell_start, const int* __restrict__ scell,
         const int* __restrict__ sidx, const double* __restrict__ sx, const double* __restrict__ sy,
         const double* __restrict__ sz, double eps2, double tol, const uint8_t* __restrict__ core_s,
         int* __restrict__ parent, unsigned long long* __restrict__ guard) {
    const int pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= m || !core_s[pos]) return;
    const int oi = sidx[pos];
    unsigned band = 0;
    for_each_candidate(G, cell_start, sx, sy, sz, scell[pos], sx[pos], sy[pos], sz[pos], [&](int j, double r) {
        if (core_s[j]) {
            if (tol > 0.0 && fabs(r - eps2) <= tol) ++band;
            if (r <= eps2) {
                const int oj = sidx[j];
                if (oj < oi) uf_union(parent, oi, oj);
            }
        }
        return true;
    });
    if (band) atomicAdd(guard, (unsigned long long)band);
}

// ---- dense grid ------------------------------------------------------------------------------------
// Neighbour cells with at least this many returns are scanned by the whole WARP for one point at a time.
constexpr int kDbTeamScan = 48;
// Tight bounding box of the points of every such HEAVY cell (fp32, rounded outwards, as order-preserving unsigned keys
// so that atomicMin / atomicMax build it): slot = (start of the cell's sorted run) / kDbTeamScan -- unique, because a
// heavy cell's run is at least that long.  A cell of a ring scan near the sensor holds a slice of a person's surface:
// its points fill a fraction of the cell's box, and two people 0.31 m apart put hundreds of returns into neighbouring
// cells that never merge.  Against the CELL box most of A survives the distance test and each survivor walks all of B;
// against B's tight box 70 % of them do not (measured on two ring frames: 285 k -> 82 k warp scans of 32 candidates).
constexpr int kDbBoxWords = 3;
__device__ __forceinline__ unsigned f32_okey(float f) {
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ double f32_okey_inv(unsigned k) {
    return (double)__uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
__global__ void __launch_bounds__(256)
db_heavy_boxes(int m, const unsigned* __restrict__ cell_start, const int* __restrict__ scell,
               const double* __restrict__ sx, const double* __restrict__ sy, const double* __restrict__ sz,
               unsigned* __restrict__ hmin, unsigned* __restrict__ hmax) {
    const int pos = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned lane = lane_id();
    const bool in = pos < m;
    const int c = in ? scell[pos] : 0;
    const unsigned s0 = in ? cell_start[c] : 0u, s1 = in ? cell_start[c + 1] : 0u;
    const bool heavy = in && (int)(s1 - s0) >= kDbTeamScan;
    const unsigned peers = __match_any_sync(0xffffffffu, heavy ? c : -1 - (int)lane);   // sorted by cell: mostly one group
    if (!heavy) return;
    const double p[3] = {sx[pos], sy[pos], sz[pos]};
    const size_t slot = (size_t)(s0 / kDbTeamScan) * kDbBoxWords;
    const bool first = (int)lane == __ffs(peers) - 1;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const unsigned lo = __reduce_min_sync(peers, f32_okey(__double2float_rd(p[a])));
        const unsigned hi = __reduce_max_sync(peers, f32_okey(__double2float_ru(p[a])));
        if (first) {
            atomicMin(hmin + slot + a, lo);
            atomicMax(hmax + slot + a, hi);
        }
    }
}
// The common case of a (point, neighbour cell) visit is "that cell has no core point" or "it is already in my
// set": both are decided from ONE load of the cell's representative (crep, filled by db_core) plus a find, without
// touching the cell's start offsets, its core flags or its index list.
template <int kAhead>
__global__ void __launch_bounds__(kDbThreads)
db_union_dense(int m, CellGrid G, const unsigned* __restrict__ cell_start, const int* __restrict__ scell,
               const int* __restrict__ sidx, const double* __restrict__ sx, const double* __restrict__ sy,
               const double* __restrict__ sz, double eps2, double tol, const uint8_t* __restrict__ core_s,
               int* __restrict__ parent, unsigned long long* __restrict__ guard, const int* __restrict__ crep,
               const unsigned* __restrict__ hmin, const unsigned* __restrict__ hmax) {
    // Every lane walks the SAME list of forward cell offsets (the union of the occupied forward cells of the warp's
    // cells), so that the warp meets at one point per offset: a lane that has to search a heavy neighbour cell (hundreds of returns near the sensor of a ring scan) does not scan it
    // alone while its 31 neighbours idle -- measured on a 128-beam frame: 2 871 heavy cell pairs never merge (two people
    // 0.31 m apart), the worst holds 770 x 619 returns, a handful of points of A survive the box test and each walked all
    // of B: 2 000 dependent candidates per thread were the kernel's 0.5 ms tail -- the warp takes the requests one by one
    // and tests 32 candidates per step, stopping at the first certain pair (lowest index, as the serial scan).
    const int pos = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned lane = lane_id();
    const bool active = pos < m && core_s[pos];
    const int oi = active ? sidx[pos] : 0;
    const int c = active ? scell[pos] : 0;
    unsigned band = 0;   // decisions of this point that rest on a pair inside the tol band (certificate)
    // (a) the core points of one cell are mutual neighbours: link to the cell's representative
    if (active) {
        // rep <= oi, and oi is usually still its own root: ONE atomic hangs it under the representative; if somebody linked
        // oi elsewhere first, that other parent and the representative are united the long way
        const int rep = crep[c];
        if (rep != oi) {
            const int old = atomicMin(&parent[oi], rep);
            if (old != oi && old != rep) uf_union(parent, old, rep);
        }
    }
    // lanes of the warp that hold core points of the same cell form a group; its first lane answers for the cell
    const unsigned peers = __match_any_sync(0xffffffffu, active ? c : -1 - (int)lane);
    const int leader = __ffs(peers) - 1;
    const bool is_leader = active && (int)lane == leader;
    int myroot = is_leader ? uf_find(parent, oi) : 0;
    // (b) neighbour cells with a larger id (the pair is examined from the smaller side): one core pair within
    //     eps merges the two cells; cells already in this point's set are skipped without a distance test
    const double x = active ? sx[pos] : 0.0, y = active ? sy[pos] : 0.0, z = active ? sz[pos] : 0.0;
    const int cz = c % G.g[2];
    const int t = c / G.g[2];
    const int cy = t % G.g[1];
    const int cx = t / G.g[1];
    constexpr int R = 2, W = 2 * R + 1;           // the dense grid's reach (make_grid); the host checks G.reach == R
    constexpr int kFwd = (W * W * W) / 2;         // 62 forward cells: (dx, dy, dz) > (0, 0, 0) lexicographically
    // squared distance from the point to the slab of cells at offset d along one axis (0 inside the own slab), shrunk
    // like cell_box_dist2: the box distance of a neighbour cell is the sum of three of these
    auto slab = [&](double p, double lo_own, int d) {
        const double lo = lo_own + d * G.cell;
        const double gap = fmax(0.0, fmax(lo - p, p - (lo + G.cell)));
        return gap * gap;
    };
    const double lox = G.min[0] + cx * G.cell, loy = G.min[1] + cy * G.cell, loz = G.min[2] + cz * G.cell;
    const double lim = (eps2 + tol + 1e-300) / (1.0 - 1e-9);        // d2 * (1 - 1e-9) - 1e-300 > eps2 + tol  <=>  d2 > lim
    // (b0) which of the 62 forward cells hold a core point at all: the first core lane of every cell of the warp looks
    //      ONCE (returns lie on surfaces: most of the 5 x 5 x 5 half-neighbourhood is empty), and the warp then walks the
    //      union of these masks instead of all 62 offsets -- every step of that walk costs the whole warp a round of
    //      shuffles and votes whether or not any lane has work in it
    unsigned long long occ = 0;
    if (is_leader) {
        int k = 0;
        for (int dx = 0; dx <= R; ++dx)
            for (int dy = (dx == 0 ? 0 : -R); dy <= R; ++dy) {
                const int ax = cx + dx, ay = cy + dy;
                const bool okxy = ax < G.g[0] && ay >= 0 && ay < G.g[1];
                const int col = okxy ? cell_id(G, ax, ay, 0) : 0;
                for (int dz = ((dx == 0 && dy == 0) ? 1 : -R); dz <= R; ++dz, ++k) {
                    const int az = cz + dz;
                    if (okxy && az >= 0 && az < G.g[2] && crep[col + az] < kNoRep) occ |= 1ull << k;
                }
            }
    }
    unsigned long long todo = ((unsigned long long)__reduce_or_sync(0xffffffffu, (unsigned)(occ >> 32)) << 32) |
                              __reduce_or_sync(0xffffffffu, (unsigned)occ);
    while (todo) {
        const int k = __ffsll((long long)todo) - 1;                  // ascending k = the lexicographic order of the plain loops
        todo &= todo - 1;
        const int Lc = kFwd + 1 + k;                                 // linear index inside the 5 x 5 x 5 cube (centre = 62)
        const int dx = Lc / (W * W) - R, dy = (Lc / W) % W - R, dz = Lc % W - R;
        // cell-level part, ONCE per cell of the warp (its first core lane): is the neighbour in this cell's set already?
        // Every core point of a cell is in the set of the cell's representative (step a), so the answer is the same for
        // all of them -- and the find()s it takes are volatile loads that go to L2
        int jf = 0, b1 = 0, c_need = 0;
        if (is_leader && ((occ >> k) & 1ull)) {
            const int nc = cell_id(G, cx + dx, cy + dy, cz + dz);
            // a root that equals the remembered root of this cell's set answers "same set" even if the remembered value
            // is old (sets only grow); a mismatch may just mean it IS old: refresh and compare again
            // -- and the walk may run on stale L1 copies: its end point is an ancestor of the neighbour's representative, so
            // meeting the remembered root proves "same set" without one coherent load.  Anything else is settled in L2, and
            // the true root is written into the walked node (path compression; the store also drops the stale L1 line)
            const int nrep = crep[nc];
            int nroot = uf_walk_cached(parent, nrep);
            if (nroot != myroot) {
                nroot = uf_find(parent, nroot);
                if (nroot != nrep && ((volatile int*)parent)[nrep] != nrep) parent[nrep] = nroot;
                myroot = uf_find(parent, myroot);
            }
            if (nroot != myroot) {
                c_need = 1;
                jf = (int)cell_start[nc];
                b1 = (int)cell_start[nc + 1];
            }
        }
        c_need = __shfl_sync(0xffffffffu, c_need, leader);
        if (!__any_sync(0xffffffffu, c_need)) continue;
        jf = __shfl_sync(0xffffffffu, jf, leader);
        b1 = __shfl_sync(0xffffffffu, b1, leader);
        // point-level part: the box of that cell must reach into this point's eps ball
        bool need = active && c_need && !(slab(x, lox, dx) + slab(y, loy, dy) + slab(z, loz, dz) > lim);
        if (need && (b1 - jf) >= kDbTeamScan) {                      // heavy neighbour: its points' own box (db_heavy_boxes)
            const size_t slot = (size_t)(jf / kDbTeamScan) * kDbBoxWords;
            const double p[3] = {x, y, z};
            double d2 = 0.0;
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const double gap = fmax(0.0, fmax(f32_okey_inv(__ldg(hmin + slot + a)) - p[a], p[a] - f32_okey_inv(__ldg(hmax + slot + a))));
                d2 += gap * gap;
            }
            need = !(d2 > lim);
        }
        // ONE pair that is certainly within eps merges the two cells, whichever lane of the cell's group finds it: the
        // group stops at the first.  Pairs inside the band are used only if the group found no certain pair, and then
        // count against the certificate (as do near misses)
        int hit = -1, maybe = -1;
        unsigned miss = 0;
        const bool heavy = need && (b1 - jf) >= kDbTeamScan;         // the same for every lane of a group
        unsigned req = __ballot_sync(0xffffffffu, heavy);
        while (req) {
            const int L = __ffs(req) - 1;
            req &= req - 1;
            const double qx = __shfl_sync(0xffffffffu, x, L), qy = __shfl_sync(0xffffffffu, y, L), qz = __shfl_sync(0xffffffffu, z, L);
            const int f0 = __shfl_sync(0xffffffffu, jf, L), f1 = __shfl_sync(0xffffffffu, b1, L);
            const unsigned group = __shfl_sync(0xffffffffu, peers, L);
            int t_hit = -1, t_maybe = -1;
            unsigned t_miss = 0;
            // one block of 32 candidates: the first certain pair in index order ends the scan; band pairs before it count
            auto judge = [&](int jbase, bool has, double cx_, double cy_, double cz_) {
                bool certain = false, in_band_in = false, in_band_out = false;
                if (has) {
                    const double r = rdist_of(qx, qy, qz, cx_, cy_, cz_);
                    if (tol > 0.0 && fabs(r - eps2) <= tol) { in_band_in = r <= eps2; in_band_out = !in_band_in; }
                    else certain = r <= eps2;
                }
                const unsigned hm = __ballot_sync(0xffffffffu, certain);
                if (tol > 0.0) {
                    const unsigned before = hm ? ((1u << (__ffs(hm) - 1)) - 1u) : 0xffffffffu;
                    const unsigned mb = __ballot_sync(0xffffffffu, in_band_in) & before;
                    t_miss += __popc(__ballot_sync(0xffffffffu, in_band_out) & before);
                    if (mb) t_maybe = jbase + 31 - __clz((int)mb);
                }
                if (hm) t_hit = jbase + __ffs(hm) - 1;
            };
            for (int j0 = f0; j0 < f1 && t_hit < 0; j0 += 32) {
                const int ja = j0 + (int)lane;
                const bool ha = ja < f1 && core_s[ja];
                double ax_ = 0, ay_ = 0, az_ = 0;
                if (ha) { ax_ = sx[ja]; ay_ = sy[ja]; az_ = sz[ja]; }
                judge(j0, ha, ax_, ay_, az_);
            }
            if ((int)lane == L) { hit = t_hit; maybe = t_maybe; miss = t_miss; }
            if (t_hit >= 0) req &= ~group;                           // the cell pair is merged: its other requests are moot
        }
        // light cells: every lane scans for its own point, kAhead (four) candidates loaded before the first is judged (the
        // scan is one dependent L2 round trip per iteration otherwise), judged in index order; the lanes of a group vote
        // after every round
        bool scanning = need && !heavy;
        int j0 = jf;
        while (__any_sync(0xffffffffu, scanning)) {
            if (scanning) {
                uint8_t cf[kAhead];
                double qx[kAhead], qy[kAhead], qz[kAhead];
#pragma unroll
                for (int q = 0; q < kAhead; ++q) {
                    const int j = j0 + q < b1 ? j0 + q : b1 - 1;
                    cf[q] = core_s[j];
                    qx[q] = sx[j]; qy[q] = sy[j]; qz[q] = sz[j];
                }
#pragma unroll
                for (int q = 0; q < kAhead; ++q) {
                    if (j0 + q >= b1 || !cf[q] || hit >= 0) continue;
                    const double r = rdist_of(x, y, z, qx[q], qy[q], qz[q]);
                    if (tol > 0.0 && fabs(r - eps2) <= tol) {
                        if (r <= eps2) maybe = j0 + q; else ++miss;
                    } else if (r <= eps2) hit = j0 + q;
                }
                j0 += kAhead;
            }
            const unsigned hm = __ballot_sync(0xffffffffu, hit >= 0);
            if (scanning && ((hm & peers) || j0 >= b1)) scanning = false;
        }
        const bool group_hit = (__ballot_sync(0xffffffffu, hit >= 0) & peers) != 0;
        if (need) {
            if (hit >= 0) uf_union(parent, oi, sidx[hit]);
            else if (!group_hit) {
                if (maybe >= 0) { uf_union(parent, oi, sidx[maybe]); ++band; }
                band += miss;
            }
        }
    }
    if (band) atomicAdd(guard, (unsigned long long)band);
}

__global__ void __launch_bounds__(kDbThreads)
db_border_dense(int m, CellGrid G, const unsigned* __restrict__ cell_start, const int* __restrict__ scell,
                const int* __restrict__ sidx, const double* __restrict__ sx, const double* __restrict__ sy,
                const double* __restrict__ sz, double eps2, double tol, const uint8_t* __restrict__ core_s,
                const int* __restrict__ label_s, int* __restrict__ labels, unsigned long long* __restrict__ guard,
                const int* __restrict__ crep) {
    const int pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= m || core_s[pos]) return;
    unsigned band = 0;
    const int c = scell[pos];
    const double x = sx[pos], y = sy[pos], z = sz[pos];
    const int cz = c % G.g[2];
    const int t = c / G.g[2];
    const int cy = t % G.g[1];
    const int cx = t / G.g[1];
    const int R = G.reach;
    const int z0 = cz - R > 0 ? cz - R : 0;
    const int z1 = cz + R < G.g[2] - 1 ? cz + R : G.g[2] - 1;
    int best = 0x7fffffff;
    for (int ax = (cx - R > 0 ? cx - R : 0); ax <= (cx + R < G.g[0] - 1 ? cx + R : G.g[0] - 1); ++ax)
        for (int ay = (cy - R > 0 ? cy - R : 0); ay <= (cy + R < G.g[1] - 1 ? cy + R : G.g[1] - 1); ++ay) {
            const int col = cell_id(G, ax, ay, 0);
            for (int az = z0; az <= z1; ++az) {
                const int rep = crep[col + az];
                if (rep >= kNoRep) continue;          // no core point in that cell
                const int lab = labels[rep];          // every core point of a cell carries the same label (written
                if (lab >= best) continue;            // by db_label_core; this kernel only writes non-core entries)
                if (cell_box_dist2(G, ax, ay, az, x, y, z) > eps2 + tol) continue;
                const int jf = (int)cell_start[col + az];
                const unsigned b1 = cell_start[col + az + 1];
                bool hit = false, maybe = false;
                unsigned miss = 0;
                for (int j0 = jf; j0 < (int)b1 && !hit; j0 += 4) {      // look-ahead of four, as in db_union_dense
                    uint8_t cf[4];
                    double qx[4], qy[4], qz[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int j = j0 + k < (int)b1 ? j0 + k : (int)b1 - 1;
                        cf[k] = core_s[j];
                        qx[k] = sx[j]; qy[k] = sy[j]; qz[k] = sz[j];
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (j0 + k >= (int)b1 || !cf[k] || hit) continue;
                        const double r = rdist_of(x, y, z, qx[k], qy[k], qz[k]);
                        if (tol > 0.0 && fabs(r - eps2) <= tol) {
                            if (r <= eps2) maybe = true; else ++miss;
                        } else if (r <= eps2) hit = true;
                    }
                }
                if (hit) best = lab;
                else {
                    if (maybe) { best = lab; ++band; }
                    band += miss;
                }
            }
        }
    labels[sidx[pos]] = best == 0x7fffffff ? -1 : best;
    if (band) atomicAdd(guard, (unsigned long long)band);
}

// ---- local point density (visualisation) -------------------------------------------------------------
// KDTree(points).query_radius(points, r, count_only=True)  (utils/visualization.py:43-45, 167-168;
// app_simplified.py:158-159): per point, the number of points (itself included) with fp64 rdist <= r*r.
// Same cell list as DBSCAN with cell edge r/2 and a 5x5x5 neighbourhood; a cell whose box lies entirely
// inside the ball is counted wholesale (the k-d tree does the same with whole nodes), entirely outside is
// skipped, the rest is tested point by point.  The box bounds are widened / shrunk by 1e-9 relative, so a
// wholesale decision is never taken on a pair that rounding could flip.
__global__ void __launch_bounds__(kDbThreads)
db_ball_count(int m, CellGrid G, const unsigned* __restrict__ cell_start, const int* __restrict__ scell,
              const int* __restrict__ sidx, const double* __restrict__ sx, const double* __restrict__ sy,
              const double* __restrict__ sz, double r2, long long* __restrict__ counts) {
    const int pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= m) return;
    const int c = scell[pos];
    const double x = sx[pos], y = sy[pos], z = sz[pos];
    const int cz = c % G.g[2];
    const int t = c / G.g[2];
    const int cy = t % G.g[1];
    const int cx = t / G.g[1];
    const int R = G.reach;
    const int z0 = cz - R > 0 ? cz - R : 0;
    const int z1 = cz + R < G.g[2] - 1 ? cz + R : G.g[2] - 1;
    long long cnt = 0;
    for (int ax = (cx - R > 0 ? cx - R : 0); ax <= (cx + R < G.g[0] - 1 ? cx + R : G.g[0] - 1); ++ax)
        for (int ay = (cy - R > 0 ? cy - R : 0); ay <= (cy + R < G.g[1] - 1 ? cy + R : G.g[1] - 1); ++ay) {
            const int col = cell_id(G, ax, ay, 0);
            unsigned b1 = cell_start[col + z0];
            for (int az = z0; az <= z1; ++az) {
                const unsigned b0 = b1;
                b1 = cell_start[col + az + 1];
                if (b0 == b1) continue;
                if (cell_box_dist2(G, ax, ay, az, x, y, z) > r2) continue;
                // farthest corner of the cell box, widened
                const double lo_x = G.min[0] + ax * G.cell, lo_y = G.min[1] + ay * G.cell, lo_z = G.min[2] + az * G.cell;
                const double fx = fmax(fabs(x - lo_x), fabs(x - (lo_x + G.cell)));
                const double fy = fmax(fabs(y - lo_y), fabs(y - (lo_y + G.cell)));
                const double fz = fmax(fabs(z - lo_z), fabs(z - (lo_z + G.cell)));
                if ((fx * fx + fy * fy + fz * fz) * (1.0 + 1e-9) <= r2) { cnt += (long long)(b1 - b0); continue; }
                for (int j = (int)b0; j < (int)b1; ++j) cnt += rdist_of(x, y, z, sx[j], sy[j], sz[j]) <= r2;
            }
        }
    counts[sidx[pos]] = cnt;
}

__global__ void db_roots(int m, const uint8_t* __restrict__ core_o, int* __restrict__ parent,
                         unsigned* __restrict__ is_root) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    unsigned r = 0;
    if (core_o[i]) {
        const int root = uf_find(parent, i);
        r = (root == i);
        if (root != i) parent[i] = root;
    }
    is_root[i] = r;
}

__global__ void db_label_core(int m, const int* __restrict__ sidx, const uint8_t* __restrict__ core_s,
                              int* __restrict__ parent, const unsigned* __restrict__ root_rank,
                              int* __restrict__ label_s, int* __restrict__ labels) {
    const int pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= m) return;
    int lab = -1;
    if (core_s[pos]) {
        const int oi = sidx[pos];
        const int root = uf_find(parent, oi);
        lab = (int)root_rank[root];
        labels[oi] = lab;
    }
    label_s[pos] = lab;
}

__global__ void __launch_bounds__(kDbThreads)
db_border(int m, CellGrid G, const unsigned* __restrict__ cell_start, const int* __restrict__ scell,
          const int* __restrict__ sidx, const double* __restrict__ sx, const double* __restrict__ sy,
          const double* __restrict__ sz, double eps2, double tol, const uint8_t* __restrict__ core_s,
          const int* __restrict__ label_s, int* __restrict__ labels, unsigned long long* __restrict__ guard) {
    const int pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= m || core_s[pos]) return;
    int best = 0x7fffffff;
    unsigned band = 0;
    for_each_candidate(G, cell_start, sx, sy, sz, scell[pos], sx[pos], sy[pos], sz[pos], [&](int j, double r) {
        if (core_s[j]) {
            if (tol > 0.0 && fabs(r - eps2) <= tol) ++band;
            if (r <= eps2) { const int l = label_s[j]; best = l < best ? l : best; }
        }
        return true;
    });
    labels[sidx[pos]] = best == 0x7fffffff ? -1 : best;
    if (band) atomicAdd(guard, (unsigned long long)band);
}

struct DbLayout {
    size_t off_cell, off_slot, off_count, off_start, off_sidx, off_scell, off_sx, off_sy, off_sz, off_parent,
        off_core_s, off_core_o, off_isroot, off_rank, off_label_s, off_hbox, off_scan, total;
};
static DbLayout db_layout(int64_t m, int64_t ncell) {
    DbLayout L;
    size_t o = 0;
    auto take = [&](size_t b) { size_t a = ws_align(o); o = a + b; return a; };
    L.off_cell = take(4 * m); L.off_slot = take(4 * m);
    L.off_count = take(4 * (ncell + 1)); L.off_start = take(4 * (ncell + 2));
    L.off_sidx = take(4 * m); L.off_scell = take(4 * m);
    L.off_sx = take(8 * m); L.off_sy = take(8 * m); L.off_sz = take(8 * m);
    L.off_parent = take(4 * m); L.off_core_s = take(m); L.off_core_o = take(m);
    L.off_isroot = take(4 * (m + 1)); L.off_rank = take(4 * (m + 2)); L.off_label_s = take(4 * m);
    L.off_hbox = take(2 * kDbBoxWords * sizeof(unsigned) * (size_t)(m / kDbTeamScan + 2));   // tight boxes of the heavy cells
    const int64_t big = m > ncell ? m : ncell;
    L.off_scan = take(scan_workspace_bytes(big + 1));
    L.total = ws_align(o);
    return L;
}

constexpr int64_t kDbMaxCells = 1ll << 24;
constexpr int64_t kDbDenseMaxCells = 1ll << 25;   // 2 x 128 MB of cell counters / starts at most

static bool make_grid(const double* mn, const double* mx, double eps, bool want_dense, CellGrid* G) {
    double cell = eps * (1.0 + 1e-9);
    if (!(cell > 0.0)) return false;
    G->reach = 1;
    G->dense = 0;
    if (want_dense) {
        // dense grid: cell diagonal < eps, neighbours two cells away; only if the directory stays small
        const double dc = eps / sqrt(3.0) * (1.0 - 1e-6);
        double prod = 1.0;
        bool ok = dc > 0.0;
        for (int c = 0; c < 3 && ok; ++c) {
            const double ext = mx[c] - mn[c];
            if (!(ext >= 0.0) || !(ext < 1e300)) return false;
            prod *= floor(ext / dc) + 1.0;
        }
        if (ok && prod <= (double)kDbDenseMaxCells) {
            G->cell = dc;
            G->reach = 2;
            G->dense = 1;
            int64_t n = 1;
            for (int c = 0; c < 3; ++c) {
                G->min[c] = mn[c];
                G->g[c] = (int)(floor((mx[c] - mn[c]) / dc) + 1.0);
                n *= G->g[c];
            }
            G->ncell = (int)n;
            return true;
        }
    }
    for (int iter = 0; iter < 64; ++iter) {
        double prod = 1.0;
        for (int c = 0; c < 3; ++c) {
            const double ext = mx[c] - mn[c];
            if (!(ext >= 0.0) || !(ext < 1e300)) return false;
            const double g = floor(ext / cell) + 1.0;
            prod *= g;
        }
        if (prod <= (double)kDbMaxCells) break;
        cell *= cbrt(prod / (double)kDbMaxCells) * 1.01;
    }
    G->cell = cell;
    int64_t n = 1;
    for (int c = 0; c < 3; ++c) {
        G->min[c] = mn[c];
        const double g = floor((mx[c] - mn[c]) / cell) + 1.0;
        if (g > 2.0e9) return false;
        G->g[c] = (int)g;
        n *= G->g[c];
    }
    if (n > kDbMaxCells * 2) return false;
    G->ncell = (int)n;
    return true;
}

// cell grid for a radius query: cell edge r/2 (reach 2) while the directory stays small, else cell >= r
static bool make_grid_radius(const double* mn, const double* mx, double r, CellGrid* G) {
    if (!(r > 0.0)) return false;
    const double dc = r * 0.5 * (1.0 + 1e-9);
    double prod = 1.0;
    for (int c = 0; c < 3; ++c) {
        const double ext = mx[c] - mn[c];
        if (!(ext >= 0.0) || !(ext < 1e300)) return false;
        prod *= floor(ext / dc) + 1.0;
    }
    if (prod > (double)kDbDenseMaxCells) return make_grid(mn, mx, r, false, G);
    G->cell = dc;
    G->reach = 2;
    G->dense = 0;
    int64_t n = 1;
    for (int c = 0; c < 3; ++c) {
        G->min[c] = mn[c];
        G->g[c] = (int)(floor((mx[c] - mn[c]) / dc) + 1.0);
        n *= G->g[c];
    }
    G->ncell = (int)n;
    return true;
}

// ------------------------------------------------------------------------------------------------
// K8 cluster centroids: exact two-limb integer sums, warp-aggregated by label
//   value = hi * 2^-20 + lo * 2^-60   (hi = rint(v * 2^20), lo = rint((v - hi*2^-20) * 2^60))
// Integer adds commute, so the result does not depend on the order in which atomics land.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
centroid_accumulate(const double* __restrict__ pts, const long long* __restrict__ labels64,
                    const int* __restrict__ labels32, int64_t n, int n_clusters,
                    long long* __restrict__ acc /*[reps][C][6]*/, unsigned* __restrict__ cnt /*[reps][C]*/, int reps) {
    // CTA b adds into replica b % reps: a person next to the sensor holds tens of thousands of returns, and every warp
    // that meets her sends seven atomics to the same two sectors -- with one copy of the accumulators those serialise
    // in L2 and the kernel waits for them (76 us for 1 M points; the sums are integers, so the fold order is free).
    acc += (size_t)(blockIdx.x % reps) * (size_t)n_clusters * 6;
    cnt += (size_t)(blockIdx.x % reps) * (size_t)n_clusters;
    const int64_t n_round = ((n + 31) / 32) * 32;
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += step) {
        int lab = -1;
        long long v[6] = {0, 0, 0, 0, 0, 0};
        if (i < n) {
            lab = labels64 ? (int)labels64[i] : labels32[i];
            if (lab >= n_clusters) lab = -1;
            if (lab >= 0) {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const double p = pts[3 * i + c];
                    const double hi = nearbyint(__dmul_rn(p, 1048576.0));
                    const double lo = __dsub_rn(p, __dmul_rn(hi, 1.0 / 1048576.0));
                    v[2 * c] = __double2ll_rn(hi);
                    v[2 * c + 1] = __double2ll_rn(__dmul_rn(lo, 1152921504606846976.0));
                }
            }
        }
        const unsigned act = __activemask();
        const unsigned peers = __match_any_sync(act, lab);
        const int leader = __ffs(peers) - 1;
        long long sum[6] = {0, 0, 0, 0, 0, 0};
        if (lab >= 0) {
            // Sum inside each peer group with REDUX, which takes an arbitrary member mask: the 64-bit addends are
            // cut into 22 + 22 + 20 bit pieces (32 of them cannot overflow 32 bits) and the piece sums are put
            // back together modulo 2^64, i.e. exactly the two's-complement sum.  (A member-by-member shuffle walk
            // cost 32 x 12 SHFL per warp on clouds in scan order, where a whole warp shares one cluster.)
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                const unsigned long long u = (unsigned long long)v[c];
                const unsigned long long s0 = __reduce_add_sync(peers, (unsigned)(u & 0x3fffffull));
                const unsigned long long s1 = __reduce_add_sync(peers, (unsigned)((u >> 22) & 0x3fffffull));
                const unsigned long long s2 = __reduce_add_sync(peers, (unsigned)(u >> 44));
                sum[c] = (long long)(s0 + (s1 << 22) + (s2 << 44));
            }
        }
        if (lab >= 0 && (int)lane_id() == leader) {
            unsigned long long* A = reinterpret_cast<unsigned long long*>(acc + (size_t)lab * 6);
#pragma unroll
            for (int c = 0; c < 6; ++c) atomicAdd(A + c, (unsigned long long)sum[c]);
            atomicAdd(cnt + lab, (unsigned)__popc(peers));
        }
    }
}

__global__ void centroid_finalize(int n_clusters, const long long* __restrict__ acc, const unsigned* __restrict__ cnt,
                                  int reps, double* __restrict__ out /*[C][3]*/, long long* __restrict__ counts) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_clusters) return;
    long long sum[6] = {0, 0, 0, 0, 0, 0};
    unsigned members = 0;
    for (int r = 0; r < reps; ++r) {                      // two's-complement integer sums: exact in any order
        const long long* A = acc + ((size_t)r * n_clusters + c) * 6;
#pragma unroll
        for (int i = 0; i < 6; ++i) sum[i] = (long long)((unsigned long long)sum[i] + (unsigned long long)A[i]);
        members += cnt[(size_t)r * n_clusters + c];
    }
    const double k = (double)members;
    for (int a = 0; a < 3; ++a) {
        const double hi = __dmul_rn((double)sum[2 * a], 1.0 / 1048576.0);
        const double lo = __dmul_rn((double)sum[2 * a + 1], 1.0 / 1152921504606846976.0);
        out[(size_t)c * 3 + a] = members ? __ddiv_rn(__dadd_rn(hi, lo), k) : 0.0;
    }
    if (counts) counts[c] = (long long)members;
}

}  // namespace lidar

using namespace lidar;

static int g_db_dense = 1;   // lidar_dbscan_set_dense: 0 forces the general grid (cross-check in the tests)

extern "C" {

int lidar_dbscan_set_dense(int on) {
    g_db_dense = on ? 1 : 0;
    return LIDAR_OK;
}

size_t lidar_dbscan_workspace_bytes(int64_t m, double eps, const double* h_min3, const double* h_max3) {
    if (m < 0 || !h_min3 || !h_max3) return 0;
    // tol is not known here: size for whichever grid is larger
    CellGrid G, Gd;
    if (!make_grid(h_min3, h_max3, eps, false, &G) || !make_grid(h_min3, h_max3, eps, true, &Gd)) return 0;
    const size_t a = db_layout(m, G.ncell).total, b = db_layout(m, Gd.ncell).total;
    return a > b ? a : b;
}

int lidar_dbscan(const double* d_points, int64_t m, double eps, int min_samples, double tol,
                 const double* h_min3, const double* h_max3, int32_t* d_labels, int32_t* d_n_clusters,
                 uint64_t* d_guard, void* d_ws, size_t ws_bytes, void* stream) {
    LIDAR_REQUIRE(m >= 0 && m < (1ll << 31) - 64, LIDAR_ERR_INVALID, "lidar_dbscan: m out of range");
    LIDAR_REQUIRE(eps > 0.0 && min_samples >= 1 && tol >= 0.0, LIDAR_ERR_INVALID, "lidar_dbscan: bad eps/min_samples/tol");
    LIDAR_REQUIRE(d_n_clusters && d_guard, LIDAR_ERR_INVALID, "lidar_dbscan: NULL output");
    cudaStream_t st = as_stream(stream);
    LIDAR_CUDA_TRY(cudaMemsetAsync(d_n_clusters, 0, sizeof(int32_t), st));
    LIDAR_CUDA_TRY(cudaMemsetAsync(d_guard, 0, sizeof(uint64_t), st));
    if (m == 0) return LIDAR_OK;
    LIDAR_REQUIRE(d_points && d_labels && h_min3 && h_max3, LIDAR_ERR_INVALID, "lidar_dbscan: NULL argument");
    CellGrid G;
    // dense grid unless the tol band could reach same-cell pairs (they sit >= 2e-6 * eps^2 below the threshold)
    LIDAR_REQUIRE(make_grid(h_min3, h_max3, eps, tol <= 1e-7 * eps * eps && g_db_dense && m < (int64_t)kNoRep, &G), LIDAR_ERR_INVALID,
                  "lidar_dbscan: cannot build a cell grid for this bbox");
    const DbLayout L = db_layout(m, G.ncell);
    LIDAR_REQUIRE(d_ws && ws_bytes >= L.total, LIDAR_ERR_WORKSPACE, "lidar_dbscan: workspace too small (%zu < %zu)",
                  ws_bytes, L.total);
    char* ws = static_cast<char*>(d_ws);
    int* cell = reinterpret_cast<int*>(ws + L.off_cell);
    int* slot = reinterpret_cast<int*>(ws + L.off_slot);
    unsigned* cell_count = reinterpret_cast<unsigned*>(ws + L.off_count);
    unsigned* cell_start = reinterpret_cast<unsigned*>(ws + L.off_start);
    int* sidx = reinterpret_cast<int*>(ws + L.off_sidx);
    int* scell = reinterpret_cast<int*>(ws + L.off_scell);
    double* sx = reinterpret_cast<double*>(ws + L.off_sx);
    double* sy = reinterpret_cast<double*>(ws + L.off_sy);
    double* sz = reinterpret_cast<double*>(ws + L.off_sz);
    int* parent = reinterpret_cast<int*>(ws + L.off_parent);
    uint8_t* core_s = reinterpret_cast<uint8_t*>(ws + L.off_core_s);
    uint8_t* core_o = reinterpret_cast<uint8_t*>(ws + L.off_core_o);
    unsigned* is_root = reinterpret_cast<unsigned*>(ws + L.off_isroot);
    unsigned* root_rank = reinterpret_cast<unsigned*>(ws + L.off_rank);
    int* label_s = reinterpret_cast<int*>(ws + L.off_label_s);
    void* scan_ws = ws + L.off_scan;
    unsigned long long* guard = reinterpret_cast<unsigned long long*>(d_guard);

    const int mi = (int)m;
    const int g256 = (mi + 255) / 256;
    const int gdb = (mi + kDbThreads - 1) / kDbThreads;
    const double eps2 = eps * eps;
    LIDAR_CUDA_TRY(cudaMemsetAsync(cell_count, 0, sizeof(unsigned) * (G.ncell + 1), st));
    db_cell_assign<<<g256, 256, 0, st>>>(d_points, mi, G, cell, slot, cell_count);
    LIDAR_CHECK_LAUNCH();
    LIDAR_CUDA_TRY(launch_exclusive_scan(cell_count, cell_start, (int64_t)G.ncell, nullptr, scan_ws, st));
    db_scatter<<<g256, 256, 0, st>>>(d_points, mi, cell, slot, cell_start, sidx, scell, sx, sy, sz, parent);
    LIDAR_CHECK_LAUNCH();
    // the per-cell counters are dead after the scan: the same array now holds the cells' representatives
    int* crep = reinterpret_cast<int*>(cell_count);
    if (G.dense) LIDAR_CUDA_TRY(cudaMemsetAsync(crep, 0x7f, sizeof(int) * G.ncell, st));
    const size_t hbox_words = (size_t)kDbBoxWords * (size_t)(m / kDbTeamScan + 2);
    unsigned* hmin = reinterpret_cast<unsigned*>(ws + L.off_hbox);
    unsigned* hmax = hmin + hbox_words;
    if (G.dense) {
        LIDAR_CUDA_TRY(cudaMemsetAsync(hmin, 0xff, sizeof(unsigned) * hbox_words, st));
        LIDAR_CUDA_TRY(cudaMemsetAsync(hmax, 0x00, sizeof(unsigned) * hbox_words, st));
        db_heavy_boxes<<<g256, 256, 0, st>>>(mi, cell_start, scell, sx, sy, sz, hmin, hmax);
        LIDAR_CHECK_LAUNCH();
    }
    db_core<<<gdb, kDbThreads, 0, st>>>(mi, G, cell_start, scell, sidx, sx, sy, sz, eps2, tol, min_samples, core_s,
                                         core_o, guard, crep);
    LIDAR_CHECK_LAUNCH();
    // look-ahead of 4 candidates (56 registers): whole DBSCAN of a 128-beam frame 0.95 -> 0.79 ms; 2: 0.80 ms, 8: 1.02 ms (92 registers)
    LIDAR_REQUIRE(!G.dense || G.reach == 2, LIDAR_ERR_INVALID, "lidar_dbscan: the dense grid kernels assume reach 2");
    if (G.dense) db_union_dense<4><<<gdb, kDbThreads, 0, st>>>(mi, G, cell_start, scell, sidx, sx, sy, sz, eps2, tol, core_s, parent, guard, crep, hmin, hmax);
    else db_union<<<gdb, kDbThreads, 0, st>>>(mi, G, cell_start, scell, sidx, sx, sy, sz, eps2, tol, core_s, parent, guard);
    LIDAR_CHECK_LAUNCH();
    db_roots<<<g256, 256, 0, st>>>(mi, core_o, parent, is_root);
    LIDAR_CHECK_LAUNCH();
    LIDAR_CUDA_TRY(launch_exclusive_scan(is_root, root_rank, (int64_t)mi, nullptr, scan_ws, st));
    LIDAR_CUDA_TRY(cudaMemcpyAsync(d_n_clusters, root_rank + mi, sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
    db_label_core<<<g256, 256, 0, st>>>(mi, sidx, core_s, parent, root_rank, label_s, d_labels);
    LIDAR_CHECK_LAUNCH();
    if (G.dense) db_border_dense<<<gdb, kDbThreads, 0, st>>>(mi, G, cell_start, scell, sidx, sx, sy, sz, eps2, tol, core_s, label_s, d_labels, guard, crep);
    else db_border<<<gdb, kDbThreads, 0, st>>>(mi, G, cell_start, scell, sidx, sx, sy, sz, eps2, tol, core_s, label_s,
                                                d_labels, guard);
    LIDAR_CHECK_LAUNCH();
    return LIDAR_OK;
}

size_t lidar_ball_count_workspace_bytes(int64_t m, double radius, const double* h_min3, const double* h_max3) {
    if (m < 0 || !h_min3 || !h_max3) return 0;
    CellGrid G;
    if (!make_grid_radius(h_min3, h_max3, radius, &G)) return 0;
    return db_layout(m, G.ncell).total;
}

int lidar_ball_count(const double* d_points, int64_t m, double radius, const double* h_min3, const double* h_max3,
                     int64_t* d_counts, void* d_ws, size_t ws_bytes, void* stream) {
    LIDAR_REQUIRE(m >= 0 && m < (1ll << 31) - 64, LIDAR_ERR_INVALID, "lidar_ball_count: m out of range");
    LIDAR_REQUIRE(radius > 0.0, LIDAR_ERR_INVALID, "lidar_ball_count: radius must be > 0");
    if (m == 0) return LIDAR_OK;
    LIDAR_REQUIRE(d_points && d_counts && h_min3 && h_max3, LIDAR_ERR_INVALID, "lidar_ball_count: NULL argument");
    CellGrid G;
    LIDAR_REQUIRE(make_grid_radius(h_min3, h_max3, radius, &G), LIDAR_ERR_INVALID,
                  "lidar_ball_count: cannot build a cell grid for this bbox");
    const DbLayout L = db_layout(m, G.ncell);
    LIDAR_REQUIRE(d_ws && ws_bytes >= L.total, LIDAR_ERR_WORKSPACE, "lidar_ball_count: workspace too small (%zu < %zu)",
                  ws_bytes, L.total);
    char* ws = static_cast<char*>(d_ws);
    int* cell = reinterpret_cast<int*>(ws + L.off_cell);
    int* slot = reinterpret_cast<int*>(ws + L.off_slot);
    unsigned* cell_count = reinterpret_cast<unsigned*>(ws + L.off_count);
    unsigned* cell_start = reinterpret_cast<unsigned*>(ws + L.off_start);
    int* sidx = reinterpret_cast<int*>(ws + L.off_sidx);
    int* scell = reinterpret_cast<int*>(ws + L.off_scell);
    double* sx = reinterpret_cast<double*>(ws + L.off_sx);
    double* sy = reinterpret_cast<double*>(ws + L.off_sy);
    double* sz = reinterpret_cast<double*>(ws + L.off_sz);
    int* parent = reinterpret_cast<int*>(ws + L.off_parent);
    void* scan_ws = ws + L.off_scan;
    cudaStream_t st = as_stream(stream);
    const int mi = (int)m;
    const int g256 = (mi + 255) / 256;
    LIDAR_CUDA_TRY(cudaMemsetAsync(cell_count, 0, sizeof(unsigned) * (G.ncell + 1), st));
    db_cell_assign<<<g256, 256, 0, st>>>(d_points, mi, G, cell, slot, cell_count);
    LIDAR_CHECK_LAUNCH();
    LIDAR_CUDA_TRY(launch_exclusive_scan(cell_count, cell_start, (int64_t)G.ncell, nullptr, scan_ws, st));
    db_scatter<<<g256, 256, 0, st>>>(d_points, mi, cell, slot, cell_start, sidx, scell, sx, sy, sz, parent);
    LIDAR_CHECK_LAUNCH();
    db_ball_count<<<(mi + kDbThreads - 1) / kDbThreads, kDbThreads, 0, st>>>(
        mi, G, cell_start, scell, sidx, sx, sy, sz, radius * radius, reinterpret_cast<long long*>(d_counts));
    LIDAR_CHECK_LAUNCH();
    return LIDAR_OK;
}

// replicas of the accumulators (centroid_accumulate): eight while that stays under a few MB, fewer for many clusters
static int centroid_reps(int n_clusters) { return n_clusters <= 8192 ? 8 : n_clusters <= 65536 ? 2 : 1; }

size_t lidar_centroid_workspace_bytes(int n_clusters) {
    if (n_clusters < 0) return 0;
    const size_t r = (size_t)centroid_reps(n_clusters);
    return ws_align(sizeof(long long) * 6 * r * (size_t)n_clusters) + ws_align(sizeof(unsigned) * r * (size_t)n_clusters);
}

int lidar_cluster_centroids(const double* d_points, const void* d_labels, int labels_are_i64, int64_t n,
                            int n_clusters, double* d_centroids3, int64_t* d_counts, void* d_ws, size_t ws_bytes,
                            void* stream) {
    LIDAR_REQUIRE(n >= 0 && n_clusters >= 0, LIDAR_ERR_INVALID, "lidar_cluster_centroids: bad sizes");
    if (n_clusters == 0) return LIDAR_OK;
    LIDAR_REQUIRE(d_centroids3 && (n == 0 || (d_points && d_labels)), LIDAR_ERR_INVALID,
                  "lidar_cluster_centroids: NULL argument");
    const size_t need = lidar_centroid_workspace_bytes(n_clusters);
    LIDAR_REQUIRE(d_ws && ws_bytes >= need, LIDAR_ERR_WORKSPACE, "lidar_cluster_centroids: workspace too small");
    cudaStream_t st = as_stream(stream);
    LIDAR_CUDA_TRY(cudaMemsetAsync(d_ws, 0, need, st));
    long long* acc = static_cast<long long*>(d_ws);
    const int reps = centroid_reps(n_clusters);
    unsigned* cnt = reinterpret_cast<unsigned*>(static_cast<char*>(d_ws) + ws_align(sizeof(long long) * 6 * (size_t)reps * (size_t)n_clusters));
    if (n > 0) {
        int64_t want = (n + 255) / 256;
        const int64_t cap = (int64_t)sm_count() * 8;
        const int grid = (int)(want < cap ? want : cap);
        centroid_accumulate<<<grid, 256, 0, st>>>(d_points, labels_are_i64 ? static_cast<const long long*>(d_labels) : nullptr,
                                                  labels_are_i64 ? nullptr : static_cast<const int*>(d_labels), n,
                                                  n_clusters, acc, cnt, reps);
        LIDAR_CHECK_LAUNCH();
    }
    centroid_finalize<<<(n_clusters + 127) / 128, 128, 0, st>>>(n_clusters, acc, cnt, reps, d_centroids3,
                                                                reinterpret_cast<long long*>(d_counts));
    LIDAR_CHECK_LAUNCH();
    return LIDAR_OK;
}

}  // extern "C"
