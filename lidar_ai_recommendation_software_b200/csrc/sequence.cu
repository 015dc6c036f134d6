// One frame of a sequence (BASELINE configs[3]) behind ONE call: variant-B preprocess (app_simplified.py:76-110: 3-sigma
// filter, 30th-percentile ground split, DBSCAN eps 0.3 / min_samples 5 on the raw non-ground points, labels scattered
// over the inliers) and the per-cluster centroids of extract_people_positions (utils/data_processing.py:251-280).
//
// The stages are the entries of preprocess.cu / dbscan.cu, unchanged; what this file adds is the HOST side between
// them -- wait for the front's descriptor, size the cell grid from its bounding box, wait for the cluster count, size the
// centroid accumulators -- in C instead of Python.  A sequence runs several frames at once on worker threads of ONE
// interpreter: the dozen Python calls of a frame (each of them a few allocations, argument marshalling, a descriptor
// unpacked into numpy) were 0.3 ms of interpreter time per frame under one GIL, a third of what the device needs for the
// frame's kernels.  The whole call runs without the GIL (ctypes releases it), waits included.
#include "common.cuh"

using namespace lidar;

extern "C" {

int lidar_sequence_frame_b(const double* d_points, int64_t n, double eps, int min_samples, double* d_inliers,
                           double* d_nonground, int32_t* d_ng_index, int32_t* d_labels, int64_t* d_full_labels,
                           double* d_centroids3, int64_t* d_counts, int centroid_cap, lidar_front_desc* d_front,
                           uint64_t* d_info2, void* h_pinned, size_t pinned_bytes, void* d_ws_front, size_t ws_front,
                           void* d_ws_dbscan, size_t ws_dbscan, void* d_ws_centroid, size_t ws_centroid,
                           lidar_sequence_frame_out* out, void* stream) {
    LIDAR_REQUIRE(out != nullptr, LIDAR_ERR_INVALID, "lidar_sequence_frame_b: NULL out");
    memset(out, 0, sizeof(*out));
    LIDAR_REQUIRE(n > 0 && d_points && d_inliers && d_nonground && d_ng_index && d_labels && d_full_labels && d_front &&
                      d_info2 && h_pinned && eps > 0.0 && min_samples > 0,
                  LIDAR_ERR_INVALID, "lidar_sequence_frame_b: bad argument");
    LIDAR_REQUIRE(pinned_bytes >= sizeof(lidar_front_desc) + 64 + (size_t)(centroid_cap > 0 ? centroid_cap : 0) * 32,
                  LIDAR_ERR_INVALID, "lidar_sequence_frame_b: page-locked staging too small");
    cudaStream_t st = as_stream(stream);
    char* hp = static_cast<char*>(h_pinned);
    // ---- everything before DBSCAN: one enqueue, one read-back ----------------------------------------------------
    int rc = lidar_preprocess_front(d_points, n, 0, d_inliers, nullptr, d_nonground, d_ng_index, nullptr, d_front, d_ws_front,
                                    ws_front, stream);
    if (rc != LIDAR_OK) return rc;
    LIDAR_CUDA_TRY(cudaMemcpyAsync(hp, d_front, sizeof(lidar_front_desc), cudaMemcpyDeviceToHost, st));
    LIDAR_CUDA_TRY(cudaStreamSynchronize(st));
    memcpy(&out->front, hp, sizeof(lidar_front_desc));
    const int64_t n_in = out->front.n_in, m = out->front.n_nonground;
    if (n_in <= 0) return LIDAR_OK;                      // the caller raises what np.percentile of an empty array raises
    // ---- clustering of the non-ground points (app_simplified.py:102-110) ------------------------------------------
    uint64_t* h_info = reinterpret_cast<uint64_t*>(hp + sizeof(lidar_front_desc));
    LIDAR_CUDA_TRY(cudaMemsetAsync(d_info2, 0, 2 * sizeof(uint64_t), st));      // lidar_dbscan writes an int32 count into the low word
    if (m > 10) {
        const double* lo = out->front.bbox_ng;
        const double* hi = out->front.bbox_ng + 3;
        const size_t need = lidar_dbscan_workspace_bytes(m, eps, lo, hi);
        LIDAR_REQUIRE(need != 0, LIDAR_ERR_INVALID, "lidar_sequence_frame_b: cannot build a cell grid for this bbox / eps");
        out->need_dbscan_ws = need;
        if (need > ws_dbscan || !d_ws_dbscan) return LIDAR_ERR_WORKSPACE;        // the caller grows the buffer and calls again
        rc = lidar_dbscan(d_nonground, m, eps, min_samples, 0.0, lo, hi, d_labels, reinterpret_cast<int32_t*>(d_info2),
                          d_info2 + 1, d_ws_dbscan, ws_dbscan, stream);
        if (rc != LIDAR_OK) return rc;
    } else {
        // fewer than 11 points: one cluster holding them all (app_simplified.py:108-110)
        if (m > 0) LIDAR_CUDA_TRY(cudaMemsetAsync(d_labels, 0, sizeof(int32_t) * (size_t)m, st));
        const uint64_t two[2] = {m > 0 ? 1ull : 0ull, 0ull};
        memcpy(h_info, two, sizeof(two));
        LIDAR_CUDA_TRY(cudaMemcpyAsync(d_info2, h_info, sizeof(two), cudaMemcpyHostToDevice, st));
    }
    rc = lidar_scatter_labels(d_labels, d_ng_index, m, d_full_labels, n_in, stream);
    if (rc != LIDAR_OK) return rc;
    LIDAR_CUDA_TRY(cudaMemcpyAsync(h_info, d_info2, 2 * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    LIDAR_CUDA_TRY(cudaStreamSynchronize(st));
    out->n_clusters = (int32_t)(h_info[0] & 0xffffffffull);
    out->guard_dbscan = h_info[1];
    // ---- centroids of the clusters (labels 0 .. n_clusters-1 by construction) ---------------------------------------
    const int nc = out->n_clusters;
    if (nc <= 0 || !d_centroids3 || !d_counts) return LIDAR_OK;
    out->need_centroid_ws = lidar_centroid_workspace_bytes(nc);
    if (nc > centroid_cap || out->need_centroid_ws > ws_centroid || !d_ws_centroid) {
        out->centroids_done = 0;                         // preprocess is complete; the caller computes the centroids itself
        return LIDAR_OK;
    }
    rc = lidar_cluster_centroids(d_inliers, d_full_labels, 1, n_in, nc, d_centroids3, d_counts, d_ws_centroid, ws_centroid, stream);
    if (rc != LIDAR_OK) return rc;
    char* hc = hp + sizeof(lidar_front_desc) + 64;
    LIDAR_CUDA_TRY(cudaMemcpyAsync(hc, d_centroids3, sizeof(double) * 3 * (size_t)nc, cudaMemcpyDeviceToHost, st));
    LIDAR_CUDA_TRY(cudaMemcpyAsync(hc + sizeof(double) * 3 * (size_t)centroid_cap, d_counts, sizeof(int64_t) * (size_t)nc,
                                   cudaMemcpyDeviceToHost, st));
    LIDAR_CUDA_TRY(cudaStreamSynchronize(st));
    out->centroids_done = 1;
    return LIDAR_OK;
}

}  // extern "C"
