/* lidar_b200 — C ABI of the B200-native point-cloud hot path.
 *
 * This is the drop-in boundary for the hot path of FortuneMU2025/LIDAR_AI_Recommendation_Software.
 * The reference is pure Python and has NO FFI; the interface each entry point replaces is therefore
 * a Python call site in the reference, cited as file:line next to every declaration.  The host-side
 * mirror of those Python callables lives in lidar_ai_recommendation_software_b200/ (ctypes over this
 * header); INTEGRATION.md shows the binding a maintainer of the reference would add.
 *
 * Rules of the ABI
 *   - plain C: pointers, sizes, scalars.  No torch / C++ types.
 *   - every pointer named d_* is DEVICE memory on the current CUDA device, h_* is HOST memory.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  All work is
 *     enqueued asynchronously on it; no entry point synchronises unless it says so.
 *   - no hidden allocation: scratch comes from the caller (`d_ws`, `ws_bytes`); each family has a
 *     *_workspace_bytes() query.  Workspaces need 256-byte alignment.
 *   - return value: 0 = LIDAR_OK, negative = error; lidar_last_error() returns a thread-local,
 *     human-readable message for the last failing call on this thread.
 *   - integer outputs are bit-exact against the CPU oracle (oracle/), see DESIGN.md.
 */
#ifndef LIDAR_B200_H_
#define LIDAR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LIDAR_ABI_VERSION 2

#define LIDAR_OK 0
#define LIDAR_ERR_INVALID (-1)   /* bad argument */
#define LIDAR_ERR_CUDA (-2)      /* a CUDA runtime call failed */
#define LIDAR_ERR_WORKSPACE (-3) /* caller workspace too small */
#define LIDAR_ERR_CAPACITY (-4)  /* an output capacity was exceeded */

/* point layouts */
#define LIDAR_FMT_F32X4 0 /* float4 x,y,z,intensity (16-byte aligned)            */
#define LIDAR_FMT_F64X3 1 /* packed (n,3) float64 rows — the reference's layout  */

/* histogram accumulation strategy (lidar_hist2d*) */
#define LIDAR_HIST_AUTO 0
#define LIDAR_HIST_GLOBAL 1 /* one RED.ADD per point straight into the L2-resident grid          */
#define LIDAR_HIST_SHARED 2 /* CTA-private shared-memory grid, warp-aggregated, flushed at the end */

/* shared-MLP implementation selector (lidar_shared_mlp_maxpool) */
#define LIDAR_MLP_AUTO 0    /* tcgen05 when the shape is 3-64-64-128 / k = 32 with fused gather, else SIMT */
#define LIDAR_MLP_SIMT 1    /* fp32 CUDA-core kernel, any widths                                           */
#define LIDAR_MLP_TCGEN05 2 /* 5th-gen tensor cores (bf16 hi/lo split operands, fp32 accumulate in TMEM)   */

const char* lidar_last_error(void);
int lidar_abi_version(void);
/* sm_count, compute capability (major*10+minor), opt-in shared memory per block of `device` */
int lidar_device_props(int device, int* sm_count, int* cc, size_t* smem_optin);

/* ------------------------------------------------------------------------------------------- *
 * K1  bounding box.   replaces np.min/np.max at utils/data_processing.py:143,207-208 and
 *     app_simplified.py:80,116-117.
 *     d_out8 = {min x,y,z,w, max x,y,z,w} as float64 (w = intensity; 0 for F64X3).  n == 0 gives
 *     +inf / -inf.  Exact (min/max are order independent).
 * ------------------------------------------------------------------------------------------- */
size_t lidar_reduce_workspace_bytes(void);
int lidar_bbox(const void* d_points, int fmt, int64_t n, double* d_out8, void* d_ws, size_t ws_bytes,
               void* stream);

/* K1b per-axis moments about a caller-supplied centre: d_out6 = {S(p-c) x,y,z, S(p-c)^2 x,y,z} in
 *     fp64 (fixed reduction tree: deterministic for a given n).  replaces np.mean / np.std at
 *     utils/data_processing.py:151-152 and app_simplified.py:88-89 (two passes: c = 0, then c = mean). */
int lidar_moments(const void* d_points, int fmt, int64_t n, const double* h_center3, double* d_out6,
                  void* d_ws, size_t ws_bytes, void* stream);

/* distance of every point from the centroid (utils/visualization.py:50-54: np.mean(points, axis=0), then
 *     np.sqrt(np.sum((points - centroid)**2, axis=1))): d_points (n,3) fp64 -> d_out double[n].  The centroid is the
 *     device moments' sum / n (a parallel sum: within 1 ulp-scale of numpy's sequential mean, rtol 1e-12). */
int lidar_centroid_distances(const double* d_points, int64_t n, double* d_out, void* d_ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------- *
 * K6  2-D histogram with numpy.histogramdd semantics: bin k <=> e[k] <= x < e[k+1], last bin
 *     closed, everything else (and NaN) dropped; edges and compares in fp64.
 *     replaces np.histogram2d at utils/data_processing.py:316-319 (np.arange edges),
 *     utils/visualization.py:130-134 and app_simplified.py:205-209 (np.linspace edges).
 *     d_u/d_v: fp64 columns with element strides su/sv (points[:,0], points[:,1] -> stride 3).
 *     d_ex (nx+1), d_ey (ny+1): monotonically increasing fp64 edges.
 *     d_counts: int32 [nx][ny] row-major (index order [x][y], NOT transposed) — the caller zeroes
 *     it; counts are ADDED so several shards can accumulate into one grid.
 * ------------------------------------------------------------------------------------------- */
int lidar_hist2d_f64(const double* d_u, int64_t su, const double* d_v, int64_t sv, int64_t n,
                     const double* d_ex, int nx, const double* d_ey, int ny, int32_t* d_counts,
                     int mode, void* stream);
/* same, x/y taken from a point cloud in layout `fmt` (F32X4 values are widened exactly) */
int lidar_hist2d_points(const void* d_points, int fmt, int64_t n, const double* d_ex, int nx,
                        const double* d_ey, int ny, int32_t* d_counts, int mode, void* stream);

/* ------------------------------------------------------------------------------------------- *
 * a5  ROI crop (NEW op): keep <=> lo <= p <= hi on x,y,z (fp32 compares
 *     for F32X4, fp64 for F64X3), order preserving.
 *     d_mask: uint8[n] (may be NULL); d_out: compacted points in the same layout (capacity n);
 *     d_count: int64 device scalar = number kept.
 * ------------------------------------------------------------------------------------------- */
size_t lidar_compact_workspace_bytes(int64_t n);
int lidar_roi_crop(const void* d_points, int fmt, int64_t n, const double* h_lo3, const double* h_hi3,
                   uint8_t* d_mask, void* d_out, int64_t* d_count, void* d_ws, size_t ws_bytes,
                   void* stream);

/* ------------------------------------------------------------------------------------------- *
 * K2-K4  per-point stages of preprocess_lidar_data (utils/data_processing.py:127-229) and
 *        preprocess_point_cloud (app_simplified.py:76-137) on (n,3) float64 clouds.
 * ------------------------------------------------------------------------------------------- */
size_t lidar_preprocess_workspace_bytes(int64_t n);

/* 3-sigma inlier filter + compaction + height colours (data_processing.py:143-157).
 *   keep <=> |p - mean| < thr on every axis (thr = 3*std, computed by the caller in fp64).
 *   colours: h = (z - zmin)/zden ; (h, 0.5*(1-h), 0.5)   (zden = zmax - zmin + 1e-10)
 *   d_guard: #points within tol of a threshold — the knife-edge certificate (0 => mask is exact even
 *   though mean/std come from a parallel reduction).
 *   d_mask uint8[n] (may be NULL); d_out_colors may be NULL. */
int lidar_sigma_filter(const double* d_points, int64_t n, const double* h_mean3, const double* h_thr3,
                       const double* h_tol3, double zmin, double zden, uint8_t* d_mask, double* d_out_points,
                       double* d_out_colors, int64_t* d_count, uint64_t* d_guard, void* d_ws, size_t ws_bytes,
                       void* stream);

/* k-th and (k+1)-th smallest of a strided fp64 column (radix select, exact): the two order statistics
 * np.percentile(z, 30) interpolates (data_processing.py:164).  d_out2 = {x_(k), x_(k+1)} (0-based;
 * x_(k+1) = x_(k) when k = n-1). */
int lidar_select_kth(const double* d_column, int64_t stride_elems, int64_t n, int64_t k, double* d_out2,
                     void* d_ws, size_t ws_bytes, void* stream);

/* ground split (data_processing.py:165-188): ground <=> z <= z_threshold.
 *   d_plane10 = {n, Sx, Sy, Sz, Sxx, Sxy, Syy, Sxz, Syz, 0} over the ground points about h_center3
 *   (normal equations of the lstsq plane z = ax + by + c);
 *   non-ground points are compacted (order preserving) into d_out_points with their row index in
 *   d_out_index (int32).  d_guard counts |z - thr| <= tol, z != thr. */
int lidar_ground_split(const double* d_points, int64_t n, double z_threshold, const double* h_center3, double tol,
                       double* d_out_points, int32_t* d_out_index, int64_t* d_count, double* d_plane10,
                       uint64_t* d_guard, void* d_ws, size_t ws_bytes, void* stream);

/* StandardScaler.transform (data_processing.py:190-191): out = (x - mean) / scale, fp64, IEEE ops */
int lidar_standardize(const double* d_points, int64_t n, const double* h_mean3, const double* h_scale3,
                      double* d_out, void* stream);

/* d_dst[i] = d_src[d_index[i]] for k rows of row_bytes bytes (multiple of 4): the gather behind
 * downsample_point_cloud (utils/data_processing.py:247-249).  Indices are not range checked. */
int lidar_gather_rows(const void* d_src, int64_t n_rows, int row_bytes, const int64_t* d_index, int64_t k,
                      void* d_dst, void* stream);

/* full_labels = -1; full_labels[index[j]] = labels[j]  (data_processing.py:203-204), int64 output */
int lidar_scatter_labels(const int32_t* d_labels, const int32_t* d_index, int64_t m, int64_t* d_full, int64_t n,
                         void* stream);

/* The stages above as ONE enqueue without host round trips (every scalar one stage hands to the next -- mean, std,
 * thresholds, the inlier count, the percentile rank and its lerp, the scaler statistics, eps -- stays on the device,
 * computed with the same float64 expressions the host path uses):
 *   bbox + mean/std of the raw cloud (data_processing.py:143,151-152) -> 3-sigma filter + colours (:153-157) ->
 *   30th-percentile of the inlier heights (:164) -> ground split + plane sums + bbox of the inliers and of the
 *   non-ground points (:165-188, 207-208) -> with LIDAR_FRONT_SCALER: StandardScaler statistics, the scaled copy and
 *   the adaptive eps (:190-196).
 * Every output array has capacity n rows; the valid row counts are d_front->n_in / n_nonground.  The caller copies
 * *d_front back once (sizeof(lidar_front_desc)) and checks n_in > 0.  key_in / key_ng are scratch. */
typedef struct lidar_front_desc {
    double bbox_raw[8];      /* lidar_bbox of the raw cloud: min x,y,z,(w) then max x,y,z,(w) */
    double sum1[6];          /* lidar_moments about 0 */
    double mean[3];
    double sum2[6];          /* lidar_moments about mean */
    double std[3], thr[3], tol[3];
    double zmin, zden;
    int64_t n_in;            /* inliers */
    uint64_t guard_sigma;
    double kth[2];           /* z_(lo), z_(lo+1) of the inliers */
    double z_thr;            /* np.percentile(z, 30) */
    int64_t n_nonground;
    uint64_t guard_ground;
    double plane[10];        /* lidar_ground_split sums about mean */
    double bbox_in[6];       /* min xyz, max xyz of the inliers */
    double bbox_ng[6];       /* ... of the non-ground points */
    double t1[6], sc_mean[3], t2[6], scale[3];   /* StandardScaler.fit of the non-ground points */
    double u1[6], xm[3], u2[6], xstd[3];         /* np.std of the scaled points */
    double eps;
    uint64_t key_in[6], key_ng[6];
} lidar_front_desc;
#define LIDAR_FRONT_COLORS 1
#define LIDAR_FRONT_SCALER 2
size_t lidar_preprocess_front_workspace_bytes(int64_t n);
int lidar_preprocess_front(const double* d_points, int64_t n, int flags, double* d_inliers, double* d_colors,
                           double* d_nonground, int32_t* d_ng_index, double* d_scaled, lidar_front_desc* d_front,
                           void* d_ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------- *
 * K7  DBSCAN with scikit-learn-identical labels (sklearn 1.9.0 DBSCAN(eps, min_samples).fit(X).labels_;
 *     data_processing.py:197, app_simplified.py:107).  d_points (m,3) fp64; h_min3/h_max3 = bbox of
 *     the points; d_labels int32[m] (-1 noise); d_n_clusters int32; d_guard = knife-edge pairs within
 *     tol of eps^2 (pass tol = 0 when X is the raw data: fp64 rdist is then exact).
 * K8  per-cluster mean (extract_people_positions, data_processing.py:251-280): exact integer sums.
 *     d_labels int32 or int64 (labels_are_i64), ids outside [0, n_clusters) are skipped;
 *     d_centroids3 (C,3) fp64; d_counts int64[C] (may be NULL).
 *     Two cell grids give the same labels: "dense" (cell diagonal < eps, 5x5x5 neighbourhood; full cells are
 *     core without a distance test and merged cells are skipped) is used when the directory fits and
 *     tol <= 1e-7 eps^2 (decisions are then taken on pairs certainly within eps; those resting on a pair
 *     inside the tol band are counted in d_guard); "general" (cell edge >= eps, every pair tested) otherwise.  lidar_dbscan_set_dense(0) forces the
 *     general grid (process-wide; used by the tests to cross-check the two).
 * ------------------------------------------------------------------------------------------- */
int lidar_dbscan_set_dense(int on);
size_t lidar_dbscan_workspace_bytes(int64_t m, double eps, const double* h_min3, const double* h_max3);
int lidar_dbscan(const double* d_points, int64_t m, double eps, int min_samples, double tol,
                 const double* h_min3, const double* h_max3, int32_t* d_labels, int32_t* d_n_clusters,
                 uint64_t* d_guard, void* d_ws, size_t ws_bytes, void* stream);
/* Local point density of the visualisation paths: sklearn KDTree(points).query_radius(points, r, count_only=True)
 * (utils/visualization.py:43-45 and 167-168, app_simplified.py:158-159): d_counts[i] = number of points j, i itself
 * included, with fp64 rdist(i,j) <= r*r.  d_points (m,3) fp64 (2-D projections: set the third column to 0);
 * h_min3/h_max3 = bbox of the points; d_counts int64[m] (KDTree returns intp). */
size_t lidar_ball_count_workspace_bytes(int64_t m, double radius, const double* h_min3, const double* h_max3);
int lidar_ball_count(const double* d_points, int64_t m, double radius, const double* h_min3, const double* h_max3,
                     int64_t* d_counts, void* d_ws, size_t ws_bytes, void* stream);
size_t lidar_centroid_workspace_bytes(int n_clusters);
int lidar_cluster_centroids(const double* d_points, const void* d_labels, int labels_are_i64, int64_t n,
                            int n_clusters, double* d_centroids3, int64_t* d_counts, void* d_ws, size_t ws_bytes,
                            void* stream);

/* ------------------------------------------------------------------------------------------- *
 * K9  lattice flow field and bottleneck scoring (models/crowd_flow_model.py:88-279,
 *     app_simplified.py:348-447).  Lattice = meshgrid(x_grid, y_grid) row-major with y outer.
 *     lidar_flow_field: unit vector to (exit_x, exit_y) rotated by sin(x*freq)*cos(y*freq)*amp, slowed
 *       inside up to 3 discs of radius 3 m (h_discs_xy = x0,y0,x1,y1,...: drawn by the caller from the
 *       legacy MT19937 stream), scaled by speed_span / max|v|; magnitudes optionally clipped.
 *       d_sums4 = {sum |v|, sum vx, sum vy, bits of max|v| before scaling}.
 *     lidar_flow_bottlenecks: variant A severity (5*gradient + 5*convergence)/2 per node, 0 if the node
 *       does not qualify (speed <= 0.5, >= 5 nodes within 3 m inclusive, >= 3 in the 3..5 m ring).
 *     lidar_flow_box_max: variant B: max speed in the open +-3 m box around nodes slower than slow_below
 *       (-1 elsewhere).
 *     lidar_radius_count: number of centres within `radius` (inclusive) of every (qx[i], qy[j]); output
 *       int32 [nqy][nqx]  (KDTree.query_radius per cell centre, app_simplified.py:266-281).
 * K10 frame-to-frame flow (NEW op): nearest previous centroid (fp32, lowest
 *     index on ties) gated at `gate`, velocity = displacement / dt; lattice field = mean velocity of
 *     the matched people within `radius` (inclusive) of each node.
 * ------------------------------------------------------------------------------------------- */
int lidar_flow_field(const double* d_xgrid, int nx, const double* d_ygrid, int ny, double exit_x, double exit_y,
                     double freq, double amp, const double* h_discs_xy, int n_discs, double speed_span, int clip,
                     double clip_lo, double clip_hi, double* d_positions, double* d_vectors, double* d_magnitudes,
                     double* d_sums4, void* stream);
int lidar_flow_bottlenecks(int nx, int ny, const double* d_positions, const double* d_vectors,
                           const double* d_magnitudes, double* d_severity, void* stream);
int lidar_flow_box_max(int nx, int ny, const double* d_positions, const double* d_magnitudes, double slow_below,
                       double* d_box_max, void* stream);
int lidar_radius_count(const double* d_centres_xy, int n_centres, const double* d_qx, int nqx, const double* d_qy,
                       int nqy, double radius, int32_t* d_counts, void* stream);
/* plot_crowd_metrics join (utils/visualization.py:306-326): for every flow lattice node the nearest density cell centre of
 * the rectilinear grid repeat(grid_x, ny) x tile(grid_y, nx) — cKDTree(centres).query(nodes, k=1) — as the flat index
 * ix*ny + iy (int64) and the Euclidean distance; with d_density_flat / d_speed also the joined density, the congestion
 * risk density / (speed + 0.1) and its maximum (the normalisation of :322-323 is risk / max * 10).  Exact fp64 squared
 * distances; an exact tie between cells goes to the lowest flat index. */
int lidar_nearest_grid_cell(const double* d_nodes_xy, int n_nodes, const double* d_gx, int nx, const double* d_gy, int ny,
                            const double* d_density_flat, const double* d_speed, int64_t* d_index, double* d_distance,
                            double* d_density_at, double* d_risk, double* d_risk_max, void* stream);
int lidar_frame_flow_match(const float* d_prev_xy, int n_prev, const float* d_cur_xy, int n_cur, float dt, float gate,
                           int32_t* d_match, float* d_velocity, void* stream);
int lidar_frame_flow_field(const double* d_lattice_xy, int n_lattice, const float* d_cur_xy, const int32_t* d_match,
                           const float* d_velocity, int n_cur, double radius, double* d_vectors, double* d_magnitudes,
                           void* stream);
/* The two above behind ONE call, with the lattice built on the device: node iy*nx + ix = (x_grid[ix], y_grid[iy]), i.e.
 * np.vstack([X.ravel(), Y.ravel()]).T of np.meshgrid(x_grid, y_grid) (models/crowd_flow_model.py:107-111).  d_velocity is
 * zeroed first (unmatched people keep 0).  The per-frame flow step of a sequence is then one staging copy in, this call,
 * one copy out. */
int lidar_frame_flow(const float* d_prev_xy, int n_prev, const float* d_cur_xy, int n_cur, float dt, float gate,
                     const double* d_x_grid, int nx, const double* d_y_grid, int ny, double radius, int32_t* d_match,
                     float* d_velocity, double* d_lattice_xy, double* d_vectors, double* d_magnitudes, void* stream);

/* ------------------------------------------------------------------------------------------- *
 * K11-K13  PointNet++-style set abstraction (NEW ops; the reference only
 *          names a classifier in windows_design.md:65 and contains no such code).
 *   lidar_fps             (B,N,3) f32 -> (B,M) int32.  idx[0] = 0, fp32 ((dx*dx+dy*dy)+dz*dz), running min
 *                         initialised to 1e10, lowest index of the maximum: bit-exact.  One thread-block
 *                         cluster per cloud (N <= 16384), points resident on chip.
 *   lidar_ball_query      first k indices (ascending) with d2 < r2 (strict, fp32); the first hit pre-fills
 *                         all k slots; no hit leaves zeros: bit-exact.  (B,M,k) int32.
 *   lidar_group_points    out[b,:3,m,j] = xyz[b,idx[b,m,j]] - new_xyz[b,m]; feature channels gathered
 *                         unchanged after them.  (B,3+C,M,k) f32, exact.
 *   lidar_gather_points   new_xyz[b,m] = xyz[b,idx[b,m]]   (the centres FPS selected)
 *   lidar_shared_mlp_maxpool   relu(W3 relu(W2 relu(W1 g + b1) + b2) + b3), max over k -> (B,c3,M) f32.
 *                         Input is either a materialised grouped tensor d_grouped (B,c_in,M,k) or the
 *                         fused gather (d_xyz, d_idx, d_new_xyz[, d_feats]) with d_grouped = NULL.
 *                         W1 (c1,c_in), W2 (c2,c1), W3 (c3,c2) row-major fp32, BatchNorm pre-folded.
 * ------------------------------------------------------------------------------------------- */
size_t lidar_fps_workspace_bytes(int b, int n);
int lidar_fps(const float* d_xyz, int b, int n, int m, int32_t* d_idx, void* d_ws, size_t ws_bytes, void* stream);
int lidar_ball_query(const float* d_xyz, const float* d_new_xyz, int b, int n, int m, float radius, int k,
                     int32_t* d_idx, void* stream);
int lidar_group_points(const float* d_xyz, const float* d_feats, const int32_t* d_idx, const float* d_new_xyz, int b,
                       int n, int m, int k, int c_feat, float* d_out, void* stream);
int lidar_gather_points(const float* d_xyz, const int32_t* d_idx, int b, int n, int m, float* d_out, void* stream);
int lidar_shared_mlp_maxpool(const float* d_xyz, const float* d_feats, const int32_t* d_idx, const float* d_new_xyz,
                             const float* d_grouped, int b, int n, int m, int k, int c_in, int c1, int c2, int c3,
                             const float* d_w1, const float* d_b1, const float* d_w2, const float* d_b2,
                             const float* d_w3, const float* d_b3, float* d_out, int impl, void* stream);

/* ------------------------------------------------------------------------------------------- *
 * K5  voxel downsample (NEW op).
 *     i_axis = floor((f64(p_axis) - origin_axis) / voxel) ; key = (ix*Dy + iy)*Dz + iz.
 *     Voxels come out in ascending key order.  All integer outputs are bit-exact and run-to-run
 *     deterministic; centroids are the fp32 rounding of an exact fixed-point mean.
 *
 *     The frame descriptor lives in DEVICE memory so a whole frame (bbox -> keys -> ranks ->
 *     centroids [+ density grid]) is enqueued without a host round trip.
 * ------------------------------------------------------------------------------------------- */
typedef struct lidar_frame_desc {
    /* inputs written by lidar_frame_prepare (device side, from the bbox) or by the host */
    double origin[4];     /* voxel origin x,y,z (+pad)                                         */
    double bbox_min[4];   /* x,y,z,intensity                                                   */
    double bbox_max[4];
    double voxel;         /* voxel edge length                                                 */
    double fix_scale_xyz; /* 2^k: fixed-point scale of (p - voxel_corner)                      */
    double fix_scale_w;   /* 2^k: fixed-point scale of intensity                               */
    int32_t dims[4];      /* Dx, Dy, Dz (+pad)                                                 */
    int64_t key_space;    /* Dx*Dy*Dz                                                          */
    /* density grid (calculate_grid_density semantics), filled when grid > 0 */
    double grid;          /* cell size g, 0 = no density grid                                  */
    double ex0, ex1, exd; /* x edges: e(0), e(1), delta   (numpy arange fill rule)             */
    double ey0, ey1, eyd;
    int32_t nx, ny;       /* bins                                                              */
    /* outputs */
    int64_t n_points;
    int64_t n_voxels;
    int32_t status;       /* 0 ok, LIDAR_ERR_CAPACITY if key_space / grid exceeded capacities  */
    int32_t fast_f32;     /* 1: origin is fp32-exact, the fp32 index guess may be used         */
    /* exact division of a key (< 2^31) by Dz and by Dy: q = (key * magic) >> (31 + shift)     */
    uint32_t magic_dz, magic_dy;
    int32_t shift_dz, shift_dy;
    /* fused kernel only — timeline of CTA 0 (ns between consecutive %globaltimer stamps):
     *  [0] load+zero+bbox  [1] barrier 1   [2] descriptor   [3] mark        [4] barrier 2
     *  [5] scan popcounts  [6] wait totals [7] prefixes     [8] barrier 3   [9] rank
     *  [10] barrier 4      [11] clean+finalize   [14] = 1 if the summary-bitmap scan ran   [15] = CTAs of the grid.
     * All zero when the five-kernel path ran.                                                 */
    uint32_t trace_ns[16];
} lidar_frame_desc;

/* one output voxel = one 32-byte sector, so a voxel is written with a single full-sector store */
typedef struct lidar_voxel {
    float x, y, z, intensity; /* centroid and mean intensity                                   */
    int32_t count;            /* member points                                                 */
    int32_t key;              /* (ix*Dy + iy)*Dz + iz                                          */
    int32_t pad[2];
} lidar_voxel;

/* capacities chosen by the caller once per stream of frames */
typedef struct lidar_frame_caps {
    int64_t max_points;
    int64_t max_key_space; /* bits of occupancy bitmap, e.g. 1<<28                             */
    int32_t max_nx, max_ny;
} lidar_frame_caps;

size_t lidar_frame_workspace_bytes(const lidar_frame_caps* caps);
/* How lidar_frame_voxel_density runs a frame.
 *   LIDAR_FRAME_FUSED        one persistent cooperative kernel (k_frame_fused): the frame is pulled
 *                            into shared memory once (TMA bulk copies) and all five phases run on
 *                            the resident points, separated by grid barriers
 *   LIDAR_FRAME_MULTIKERNEL  five dependent kernels (prep, mark, scan, rank, finalize)
 *   LIDAR_FRAME_AUTO         fused, falling back to the five kernels if the cooperative launch is
 *                            refused
 * threads / ctas_per_sm / smem_kb tune the fused kernel (0 keeps the current value; smem_kb 0 =
 * as much as the frame needs).  Both paths produce identical outputs.  Process-wide setting. */
enum { LIDAR_FRAME_AUTO = 0, LIDAR_FRAME_MULTIKERNEL = 1, LIDAR_FRAME_FUSED = 2, LIDAR_FRAME_PARTITIONED = 3 };
int lidar_frame_set_fused(int mode, int threads, int ctas_per_sm, int smem_kb);
/*   LIDAR_FRAME_PARTITIONED  one persistent kernel (k_frame_part) built around an MSD radix partition by voxel-key
 *                            range: points are binned into <= 2048 partitions of 2^18 voxel cells, 8-byte {key, cell}
 *                            entries travel to the partition's owner CTA in run-coalesced writes, and the occupancy bits,
 *                            their popcount prefix and the density cells of the owner's slab live in shared memory: no
 *                            occupancy bitmap in global memory.  Needs caps.max_key_space <= 2^29 and the frame resident
 *                            in shared memory (<= ~1.3 M points on 148 SMs), else LIDAR_ERR_CAPACITY.
 * lidar_frame_set_partition_auto(1): LIDAR_FRAME_AUTO takes the partitioned back end whenever a frame is eligible (and the
 * scan-order variant is not selected), the fused one otherwise. */
int lidar_frame_set_partition_auto(int on);
/* EXPERIMENT knob, off by default: launch k_frame_fused with an ordinary launch instead of a
 * cooperative one.  Cooperative launches of different streams do not overlap on the device; ordinary
 * ones do, but then co-residency of the grid is the caller's responsibility: the grid must fit the
 * device next to whatever else is running (the kernel traps after spinning for two seconds). */
int lidar_frame_set_fused_plain_launch(int on);
/* Programmatic dependent launch of k_frame_fused (off by default): when frames are enqueued back to back on
 * one stream, the next frame's kernel may start on SMs the current one has left; everything that touches the
 * pipeline's workspace or outputs waits (griddepcontrol.wait) until the previous kernel has completed.
 *   on = 1  safe for any producer: the frame itself is read after griddepcontrol.wait too (a frame written by the
 *           kernel just before this one on the stream -- a crop, a concatenation -- is visible only then); what
 *           overlaps is the launch latency and the CTA set-up
 *   on = 2  the caller vouches that the frame was COMPLETE in device memory before the launch was enqueued (resident
 *           frames, or an event wait after the producer): the TMA load and the bounding box, which touch only the
 *           input, additionally run under the previous frame's tail
 * Outputs are identical.  Process-wide setting. */
int lidar_frame_set_fused_pdl(int on);
/* Keep up to `bytes` of the occupancy groups resident in L2 (access policy window on every k_frame_fused launch,
 * persisting hits / streaming misses; cudaLimitPersistingL2CacheSize is raised to match; clamped to what the
 * device allows).  0 = off (default).  Process-wide. */
int lidar_frame_set_fused_l2_persist(size_t bytes);
/* The scan-order variant of k_frame_fused (off by default; a per-host-thread setting, so that pipelines driven from
 * different threads choose independently): for frames as a sensor delivers them
 * (adjacent points adjacent in space) and / or key spaces much larger than the data (a 240 m x 240 m ring scan).
 *   - run-length aggregation across adjacent lanes: only the first lane of a run of equal voxel keys touches the
 *     occupancy bitmap, only the first lane of a run of equal density cells issues the reduction (with the run
 *     length), runs of later members of one voxel are summed with a segmented warp scan before they reach the
 *     accumulators: most same-address L2 atomics of such a frame disappear;
 *   - when there are more than 1.5 occupancy groups per point, scan and clean walk a summary bitmap of the occupied
 *     groups instead of streaming all of them: the frame costs what its data costs, not what its bounding box costs.
 * Outputs are identical to the default variant (integer sums, same order).  The default variant is the faster one
 * on shuffled, dense frames (the benchmark): the kernel is instruction-cache bound and does not carry this code. */
int lidar_frame_set_fused_scan_order(int on);
/* diagnostics: byte offset inside the workspace of uint64 stamps[ctas][16] (%globaltimer, ns) that
 * every CTA of the last k_frame_fused launch wrote at its 13 trace points (see trace_ns). */
size_t lidar_frame_trace_offset(const lidar_frame_caps* caps);
/* tuning knob: cap of the per-point frame kernels' grids in CTAs per SM (1..8, default 8).  Smaller
 * grids leave room for the kernels of other frames (other streams) to run concurrently. */
int lidar_frame_set_ctas_per_sm(int ctas_per_sm);
/* one-time (or after an error): zero the persistent parts of the workspace */
int lidar_frame_workspace_init(void* d_ws, size_t ws_bytes, const lidar_frame_caps* caps, void* stream);

/* whole frame: bbox -> voxelise (+ density histogram when grid_size > 0).
 *   d_points      float4[n]
 *   h_origin3     NULL => per-axis min of the cloud (B.1 default)
 *   h_xy_range4   NULL => (xmin,xmax,ymin,ymax) of the cloud; else the caller's x_range/y_range
 *                 (calculate_grid_density(positions, x_range, y_range, g),
 *                  utils/data_processing.py:282-328)
 *   d_voxel_key   int32[n]   per-point voxel key (B.1 voxel_idx)
 *   d_inverse     int32[n]   rank of the point's voxel
 *   d_voxels      lidar_voxel[cap n]  centroid (x,y,z,mean intensity), count and key per voxel,
 *                 ascending key
 *   d_grid        int32[max_nx*max_ny] laid out [nx][ny] with the ACTUAL ny from the descriptor
 *   d_desc        device lidar_frame_desc (read it back after the stream is synchronised)
 */
int lidar_frame_voxel_density(const void* d_points, int64_t n, double voxel_size, double grid_size,
                              const double* h_origin3, const double* h_xy_range4, int32_t* d_voxel_key,
                              int32_t* d_inverse, lidar_voxel* d_voxels, int32_t* d_grid, lidar_frame_desc* d_desc,
                              const lidar_frame_caps* caps, void* d_ws, size_t ws_bytes, void* stream);

/* Repack the first min(n_voxels, capacity) voxel records as structure-of-arrays for the trip to the
 * host: d_centroids4 float[capacity*4] (x,y,z,mean intensity), d_counts int32[capacity], d_keys
 * int32[capacity] or NULL.  n_voxels is read from the DEVICE descriptor (no host round trip).  This
 * is the layout the numpy surface returns (centroids (V,4), counts (V,)). */
int lidar_frame_pack_soa(const lidar_voxel* d_voxels, const lidar_frame_desc* d_desc, int64_t capacity,
                         float* d_centroids4, int32_t* d_counts, int32_t* d_keys, void* stream);

/* Same call, additionally recording six caller-created cudaEvent_t (passed as void*) on `stream`:
 * before k_frame_prep, then after each of prep, mark, scan, rank, finalize — so a benchmark can
 * attribute device time to each kernel without a profiler. */
int lidar_frame_voxel_density_timed(const void* d_points, int64_t n, double voxel_size, double grid_size,
                                    const double* h_origin3, const double* h_xy_range4, int32_t* d_voxel_key,
                                    int32_t* d_inverse, lidar_voxel* d_voxels, int32_t* d_grid,
                                    lidar_frame_desc* d_desc, const lidar_frame_caps* caps, void* d_ws,
                                    size_t ws_bytes, void* stream, void** h_events6);

/* ------------------------------------------------------------------------------------------- *
 * K5, general path: voxel downsample with 64-bit keys by a deterministic radix sort by voxel key + segmented
 * reduction ("int64; use 32-bit when it fits").  No key-space limit: a far outlier or a
 * 1 km x 1 km x 30 m venue at 0.05 m (2.4e11 cells) is fine; the occupancy-bitmap frame kernels above stop at 2^31.
 * Optional ROI crop fused in front (Appendix B.2: keep <=> lo <= p <= hi on x,y,z, fp32 compares against the bounds
 * cast to fp32): cropped points get voxel_key = inverse = -1 and take no part in the bbox / origin.
 *   d_voxel_key   int64[n]   (ix*Dy + iy)*Dz + iz per point            d_inverse  int32[n] rank of the point's voxel
 *   d_centroids4  float[4 n] x,y,z,mean intensity per voxel (capacity n), ascending key
 *   d_counts      int32[n]   d_unique_keys int64[n]                     d_desc     origin, dims, counts, status
 * One enqueue, no host round trip: the number of 8-bit sort passes follows the key width found on the device (the
 * launches of the passes a frame does not need return at once).
 * ------------------------------------------------------------------------------------------- */
typedef struct lidar_sorted_desc {
    double origin[3], bbox_min[3], bbox_max[3];
    double voxel, fix_scale_xyz, fix_scale_w;
    int64_t dims[3];
    int64_t key_space;     /* Dx*Dy*Dz (< 2^63) */
    int64_t n_points;      /* input points */
    int64_t n_kept;        /* points inside the ROI box (= n_points without one) */
    int64_t n_voxels;
    int32_t passes;        /* radix passes that ran */
    int32_t status;        /* 0, LIDAR_ERR_INVALID (a point below a caller-given origin), LIDAR_ERR_CAPACITY (key space >= 2^63) */
} lidar_sorted_desc;
size_t lidar_voxel_sorted_workspace_bytes(int64_t n);
int lidar_voxel_downsample_sorted(const void* d_points, int64_t n, double voxel_size, const double* h_origin3,
                                  const double* h_roi_lo3, const double* h_roi_hi3, int64_t* d_voxel_key, int32_t* d_inverse,
                                  float* d_centroids4, int32_t* d_counts, int64_t* d_unique_keys, lidar_sorted_desc* d_desc,
                                  void* d_ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------- *
 * The same frame from HOST buffers, in ONE call (the path a sensor driver or the numpy surface
 * takes): copy-in, the frame kernel(s), the SoA repack and ONE copy-out are enqueued on `stream`.
 *
 *   h_points      (n,4) float32 in page-locked host memory
 *   d_points      device staging for the frame, >= 16*n bytes
 *   d_voxels      lidar_voxel[n] scratch (the records before the repack)
 *   d_out, h_out  result block, device and page-locked host copy, >= lidar_frame_host_block_bytes()
 *
 * Result block layout for a frame of n points (np = n rounded up to a multiple of 8, every array
 * 32-byte aligned); lidar_frame_host_block_layout() returns the byte offsets:
 *   [0] voxel_key int32[np] | [1] inverse int32[np] | [2] centroids float[np*4] | [3] counts int32[np]
 *   | [4] unique keys int32[np] | [5] grid int32[max_nx*max_ny] | [6] lidar_frame_desc
 * Per-voxel arrays hold n_voxels (desc) valid rows.  The block is copied at the size of THIS frame,
 * so a short frame costs proportionally fewer PCIe bytes.  flags: LIDAR_HOST_UNIQUE_KEYS adds array [4]
 * (otherwise empty); LIDAR_HOST_NO_PER_POINT leaves voxel_key / inverse on the device (the copy starts at
 * the centroids).  Nothing synchronises: wait on the stream (or an event recorded after the call) before
 * reading h_out.
 * ------------------------------------------------------------------------------------------- */
enum { LIDAR_HOST_UNIQUE_KEYS = 1, LIDAR_HOST_NO_PER_POINT = 2 };
size_t lidar_frame_host_block_bytes(int64_t n, const lidar_frame_caps* caps, int flags);
int lidar_frame_host_block_layout(int64_t n, const lidar_frame_caps* caps, int flags, size_t* h_offsets7);
int lidar_frame_voxel_density_host(const void* h_points, int64_t n, double voxel_size, double grid_size,
                                   const double* h_origin3, const double* h_xy_range4, void* d_points,
                                   lidar_voxel* d_voxels, void* d_out, void* h_out, int flags,
                                   const lidar_frame_caps* caps, void* d_ws, size_t ws_bytes, void* stream);
/* Two-stage read-back: _begin enqueues copy-in, the frame and the repack but copies only the DESCRIPTOR back;
 * once the host has read n_voxels / nx / ny from it, lidar_frame_host_fetch enqueues copies of exactly what the
 * frame produced (8n bytes of per-point outputs unless LIDAR_HOST_NO_PER_POINT, 20 bytes per VOXEL, 4*nx*ny bytes
 * of grid) into the same block layout.  A sensor frame with 0.2 voxels per point moves 60 % fewer bytes than the
 * one-copy form, which has to size the per-voxel arrays by the frame. */
int lidar_frame_voxel_density_host_begin(const void* h_points, int64_t n, double voxel_size, double grid_size,
                                         const double* h_origin3, const double* h_xy_range4, void* d_points,
                                         lidar_voxel* d_voxels, void* d_out, void* h_out, int flags,
                                         const lidar_frame_caps* caps, void* d_ws, size_t ws_bytes, void* stream);
int lidar_frame_host_fetch(int64_t n, int64_t n_voxels, int nx, int ny, const void* d_out, void* h_out, int flags,
                           const lidar_frame_caps* caps, void* stream);


/* ------------------------------------------------------------------------------------------- *
 * Point-sharded density grid of one oversized scan (BASELINE configs[4], SURVEY.md 8e "points"):
 * calculate_grid_density (utils/data_processing.py:282-328) of a scan whose points are spread over the
 * GPUs of one box.  Every rank bins its shard into the SAME np.arange edges, the integer grids are
 * summed (bit-identical to the single-GPU grid), density = counts / g^2.
 *
 * lidar_scan_density: the whole call as ONE persistent cooperative kernel per rank -- local bbox, bbox
 *   exchange with the peers, arange parameters derived on the device, histogram, two-shot all-reduce of
 *   the grid over NVLink (multimem.ld_reduce.add.u32 + multimem.st through the NVSwitch when the
 *   symmetric buffer has a multicast mapping, peer loads / stores otherwise), density and cell centres.
 *   No host round trip, no NCCL call.  `epoch` > 0 numbers the calls made with this workspace / communicator
 *   (+1 per call; the peers' flags and the host-mapped descriptor carry it).  All ranks of `comm` must call it
 *   with the same epoch, grid_size and capacities; their kernels wait for each other (one rank per GPU: never
 *   two ranks on one GPU).
 *     d_grid         int32[cap_cells] grid of this rank when comm is NULL / world 1 (with a communicator
 *                    the grid lives in the symmetric buffer at lidar_scan_symm_grid_offset())
 *     d_density      double[cap_cells] -> [nx][ny] valid;  d_gx double[max_nx], d_gy double[max_ny]
 *     d_desc         device descriptor;  h_desc_mapped: optional host-MAPPED copy (cudaHostAllocMapped /
 *                    lidar_host_alloc) that the kernel fills as soon as the edges are known, with
 *                    .pad = the call's epoch -- poll it to enqueue an exactly-sized read-back behind the kernel
 *   status: 0 ok, LIDAR_SCAN_EMPTY (no point on any rank: the reference returns (None, None, None),
 *   data_processing.py:297-298), LIDAR_ERR_CAPACITY (nx > max_nx, ny > max_ny or nx*ny > cap_cells).
 *
 * lidar_scan_bbox_packed / lidar_scan_hist / lidar_scan_finish: the same phases as three enqueues for a
 *   caller that supplies the two collectives itself between them (lidar_nccl_allreduce below, or any other):
 *   MAX over d_packed4 = {-minx, -miny, maxx, maxy}, SUM over the int32 grid.  lidar_scan_hist derives the
 *   descriptor from the reduced d_packed4 on the device, zeroes the grid and bins the shard.
 * ------------------------------------------------------------------------------------------- */
#define LIDAR_SCAN_EMPTY 1
typedef struct lidar_scan_desc {
    double bbox[4];        /* global minx, miny, maxx, maxy                                        */
    double grid;           /* cell size g                                                          */
    double ex0, ex1, exd;  /* x edges: e(0), e(1), delta (numpy arange fill rule, Appendix A.2)    */
    double ey0, ey1, eyd;
    int64_t n_local;       /* points of this rank                                                  */
    int32_t nx, ny;        /* bins                                                                 */
    int32_t status;
    int32_t pad;           /* mapped host copy: the epoch of the call that wrote it                */
} lidar_scan_desc;
typedef struct lidar_scan_comm {
    int32_t rank, world;      /* world <= 16                                                        */
    size_t symm_bytes;        /* size of every rank's symmetric buffer, >= lidar_scan_symm_bytes()  */
    void* peer_ptrs[16];      /* rank r's buffer as mapped into THIS process (peer_ptrs[rank] = own) */
    void* multicast_ptr;      /* multicast mapping of the same buffer, or NULL                      */
} lidar_scan_comm;
size_t lidar_scan_workspace_bytes(void);
int lidar_scan_workspace_init(void* d_ws, size_t ws_bytes, void* stream);   /* once: zero it */
size_t lidar_scan_symm_bytes(int64_t cap_cells);   /* flags + bbox slots + grid; zero it once, before the first call */
size_t lidar_scan_symm_grid_offset(void);
int lidar_scan_density(const void* d_points, int fmt, int64_t n, double grid_size, int max_nx, int max_ny,
                       int64_t cap_cells, int32_t* d_grid, double* d_density, double* d_gx, double* d_gy,
                       lidar_scan_desc* d_desc, lidar_scan_desc* h_desc_mapped, const lidar_scan_comm* comm,
                       uint32_t epoch, void* d_ws, size_t ws_bytes, void* stream);
int lidar_scan_bbox_packed(const void* d_points, int fmt, int64_t n, double* d_packed4, void* d_ws, size_t ws_bytes,
                           void* stream);
int lidar_scan_hist(const void* d_points, int fmt, int64_t n, const double* d_packed4, double grid_size, int max_nx,
                    int max_ny, int64_t cap_cells, int32_t* d_grid, lidar_scan_desc* d_desc, void* stream);
int lidar_scan_finish(const int32_t* d_grid, const lidar_scan_desc* d_desc, double* d_density, double* d_gx,
                      double* d_gy, void* stream);

/* ------------------------------------------------------------------------------------------- *
 * NCCL plumbing for callers without torch.distributed (and the checked fallback of lidar_scan_density):
 * libnccl.so.2 is resolved at run time (dlopen), so the core library carries no link-time dependency.
 *   lidar_nccl_unique_id   rank 0 creates the 128-byte id and ships it to the other ranks by any means
 *   lidar_nccl_comm_init   collective: every rank of the job, after cudaSetDevice
 *   lidar_nccl_allreduce   in place on `stream`; op LIDAR_NCCL_MAX_F64 (the packed bbox) or
 *                          LIDAR_NCCL_SUM_I32 (the density grid); integer sums and max are order independent
 * ------------------------------------------------------------------------------------------- */
enum { LIDAR_NCCL_SUM_I32 = 0, LIDAR_NCCL_MAX_F64 = 1 };
int lidar_nccl_available(void);
int lidar_nccl_unique_id(void* h_id128);
int lidar_nccl_comm_init(const void* h_id128, int rank, int world, void** comm_out);
int lidar_nccl_comm_destroy(void* comm);
int lidar_nccl_allreduce(void* comm, void* d_buf, int64_t count, int op, void* stream);

/* ------------------------------------------------------------------------------------------- *
 * One frame of a sequence (BASELINE configs[3]) behind one call: lidar_preprocess_front (no colours, no scaler) -> wait for
 * its descriptor -> lidar_dbscan(eps, min_samples, tol 0) on the non-ground points inside the descriptor's bounding box
 * (m <= 10: one cluster, app_simplified.py:108-110) -> lidar_scatter_labels over the inliers -> wait for the cluster count
 * -> lidar_cluster_centroids of the inliers (utils/data_processing.py:251-280) -> wait.  Same entries, same results as
 * calling them one by one; the host steps between them run in C, so a worker thread of a Python process spends the frame
 * outside the interpreter lock.  Capacities: every per-point buffer holds n rows; d_centroids3 / d_counts hold
 * centroid_cap clusters; h_pinned = page-locked staging of at least sizeof(lidar_front_desc) + 64 + 32 * centroid_cap
 * bytes: [descriptor | n_clusters, guard | centroids (centroid_cap x 3 fp64) | counts (centroid_cap int64)].
 * Returns LIDAR_ERR_WORKSPACE with out->need_dbscan_ws set when the DBSCAN workspace is too small for this frame's bounding
 * box (grow it and call again).  out->centroids_done = 0 when the frame has more clusters than centroid_cap or the centroid
 * workspace is too small (out->need_centroid_ws): everything else is complete, compute the centroids separately.
 * ------------------------------------------------------------------------------------------- */
typedef struct lidar_sequence_frame_out {
    lidar_front_desc front;
    int32_t n_clusters;
    int32_t centroids_done;
    uint64_t guard_dbscan;
    uint64_t need_dbscan_ws;
    uint64_t need_centroid_ws;
} lidar_sequence_frame_out;
int lidar_sequence_frame_b(const double* d_points, int64_t n, double eps, int min_samples, double* d_inliers,
                           double* d_nonground, int32_t* d_ng_index, int32_t* d_labels, int64_t* d_full_labels,
                           double* d_centroids3, int64_t* d_counts, int centroid_cap, lidar_front_desc* d_front,
                           uint64_t* d_info2, void* h_pinned, size_t pinned_bytes, void* d_ws_front, size_t ws_front,
                           void* d_ws_dbscan, size_t ws_dbscan, void* d_ws_centroid, size_t ws_centroid,
                           lidar_sequence_frame_out* out, void* stream);

/* ------------------------------------------------------------------------------------------- *
 * Host side of the copies (the drop-in surface takes and returns numpy arrays, SURVEY.md 8b "Ownership").
 *   lidar_bind_to_device_numa   pin the calling thread (and the threads it creates) to the CPUs of the NUMA
 *       node `device` hangs off and prefer that node's memory (sched_setaffinity + set_mempolicy from
 *       /sys/bus/pci/devices/<bdf>/{numa_node,local_cpulist}); call it BEFORE allocating page-locked
 *       buffers.  Returns the node in *node_out (-1: the platform reports none; nothing is changed).
 *   lidar_host_alloc / free     page-locked host memory (cudaHostAlloc, portable + mapped), first-touched by
 *       the calling thread so the pages land on its NUMA node.
 *   lidar_host_copy_threads     size of the worker pool behind lidar_host_memcpy (default: min(8, CPUs of
 *       the calling thread's affinity mask / 2)).
 *   lidar_host_memcpy           memcpy split over the worker pool: pageable numpy <-> page-locked staging at
 *       the memory system's rate instead of one core's.
 *   lidar_host_copy_nontemporal 1: slices of 256 KB and more are written with non-temporal stores (no
 *       read-for-ownership of staging memory the CPU never reads back); 0: plain memcpy.
 *   lidar_copy_async            the DMA leg between page-locked staging and the device.
 * ------------------------------------------------------------------------------------------- */
int lidar_bind_to_device_numa(int device, int* node_out, int* ncpus_out);
int lidar_host_alloc(size_t bytes, void** h_ptr_out);
int lidar_host_free(void* h_ptr);
int lidar_host_copy_threads(int threads);
int lidar_host_memcpy(void* dst, const void* src, size_t bytes);
int lidar_host_copy_nontemporal(int on);
/* several copies as ONE job for the pool (the arrays of a frame's result): the segments are treated as one byte range
 * and cut into equal slices, so the workers are woken once.  count <= 64. */
int lidar_host_memcpy_batch(int count, void* const* dst, const void* const* src, const size_t* bytes);
/* wake the copy workers now: they spin (up to ~1 ms) for the next job instead of being woken through the futex when it
 * arrives.  Call it right after enqueuing the DMA whose completion the copy waits for. */
int lidar_host_copy_wake(void);
/* cudaMemcpyAsync between page-locked host memory and the device on `stream` (to_device: host -> device) */
int lidar_copy_async(void* dst, const void* src, size_t bytes, int to_device, void* stream);
/* cudaStreamSynchronize(stream): the wait of the small read-backs, callable from the binding without a stream object. */
int lidar_stream_synchronize(void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LIDAR_B200_H_ */
