// numpy edge arithmetic shared by the frame kernels (voxel.cu) and the point-sharded scan (scan.cu).
//
// calculate_grid_density (utils/data_processing.py:305-313) builds its edges with
//   np.arange(min - 2g, (max + 2g) + g, g)
// and numpy fills an arange from its first two elements:
//   e(0) = a, e(1) = fl(a + g), e(i) = fl(a + fl(i * delta)) with delta = fl(e(1) - a).
// Every operation is a separately rounded IEEE double operation (__dadd_rn / __dmul_rn), never an FMA.
#pragma once
#include "common.cuh"

namespace lidar {

struct ArangeAxis {
    double a, e1, delta;   // e(0), e(1), fill step
    int nb;                // bins = edges - 1
    int status;            // 0, or LIDAR_ERR_CAPACITY when the axis has no bin / more than max_bins
};

// Edges of one axis of calculate_grid_density for the data range [lo, hi] and cell size g.
// `have_points` = false (empty cloud) suppresses the "no bin" error: the caller reports the empty case.
__device__ __forceinline__ ArangeAxis arange_axis(double lo, double hi, double g, int max_bins, bool have_points) {
    ArangeAxis A;
    const double margin = __dmul_rn(g, 2.0);
    A.a = __dsub_rn(lo, margin);
    const double stop = __dadd_rn(__dadd_rn(hi, margin), g);
    const double len = ceil(__ddiv_rn(__dsub_rn(stop, A.a), g));
    const int nedges = (len > 0.0 && len < 1.0e9) ? (int)len : 0;
    A.e1 = __dadd_rn(A.a, g);
    A.delta = __dsub_rn(A.e1, A.a);
    A.nb = nedges - 1;
    A.status = 0;
    if (A.nb < 1) { A.nb = 0; if (have_points) A.status = LIDAR_ERR_CAPACITY; }
    if (A.nb > max_bins) A.status = A.status ? A.status : LIDAR_ERR_CAPACITY;
    return A;
}

// analytic arange edge (DOUBLE_fill): e(0)=a, e(1)=fl(a+g), e(i)=fl(a + fl(i*delta))
__device__ __forceinline__ double arange_edge(double a, double e1, double d, int i) {
    return i == 0 ? a : (i == 1 ? e1 : __dadd_rn(a, __dmul_rn((double)i, d)));
}
// np.histogramdd bin of x on those edges (Appendix A.1): e(k) <= x < e(k+1), last bin closed, -1 outside / NaN
__device__ __forceinline__ int arange_bin(double x, double a, double e1, double d, double rd, int nb) {
    const double hi = arange_edge(a, e1, d, nb);
    if (!(x >= a) || !(x <= hi)) return -1;
    if (x == hi) return nb - 1;
    int k = (int)floor(__dmul_rn(__dsub_rn(x, a), rd));   // guess; corrected against the exact edges
    k = k < 0 ? 0 : (k > nb - 1 ? nb - 1 : k);
    while (x < arange_edge(a, e1, d, k)) --k;
    while (x >= arange_edge(a, e1, d, k + 1)) ++k;
    return k;
}
// fp32 guess of the histogram bin, VERIFIED against the exact fp64 edges; falls back to arange_bin
__device__ __forceinline__ int fast_arange_bin(float xf, double x, float af, float rdf, double a, double e1, double d,
                                               double rd, int nb) {
    int k = (int)floorf(__fmul_rn(__fsub_rn(xf, af), rdf));
    k = k < 0 ? 0 : (k > nb - 1 ? nb - 1 : k);
    double lo = __dadd_rn(a, __dmul_rn((double)k, d));
    double hi = __dadd_rn(a, __dmul_rn((double)(k + 1), d));
    lo = k == 1 ? e1 : lo;
    hi = k == 0 ? e1 : hi;
    if (x >= lo && x < hi) return k;
    return arange_bin(x, a, e1, d, rd, nb);
}

}  // namespace lidar
