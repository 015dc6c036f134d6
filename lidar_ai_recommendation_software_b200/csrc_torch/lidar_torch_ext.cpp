// The thin PyTorch C++ extension over the C-ABI core (BASELINE.json north star: "called from Python through a thin
// PyTorch C++ extension with a C-ABI core").
//
// Nothing is computed here: every operator checks its tensors, takes the CURRENT CUDA stream of the tensor's device
// from PyTorch (c10::cuda::getCurrentCUDAStream — no `int(stream.cuda_stream)` round trip through Python), calls one
// extern "C" entry of liblidar_b200.so (include/lidar_b200.h) and RETURNS its status (0 = ok); the Python side turns a
// negative status into the package's LidarError with lidar_last_error()'s text, exactly as it does for the ctypes
// binding (no C++ exception crosses the dispatcher for a C-ABI status).  Registered as torch.ops.lidar_b200.* (TORCH_LIBRARY:
// no Python.h, no pybind).  The ctypes binding (_capi.py) stays the complete one; these are the per-frame hot calls,
// where the marshalling of 15 ctypes arguments costs more than the launch itself.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/library.h>

#include "lidar_b200.h"

namespace {

void* stream_of(const at::Tensor& t) { return at::cuda::getCurrentCUDAStream(t.get_device()).stream(); }

void need_cuda(const at::Tensor& t, const char* name) {
    TORCH_CHECK(t.is_cuda() && t.is_contiguous(), name, " must be a contiguous CUDA tensor");
}
int point_format(const at::Tensor& p) {
    need_cuda(p, "points");
    if (p.dim() == 2 && p.size(1) == 4 && p.scalar_type() == at::kFloat) return LIDAR_FMT_F32X4;
    if (p.dim() == 2 && p.size(1) == 3 && p.scalar_type() == at::kDouble) return LIDAR_FMT_F64X3;
    TORCH_CHECK(false, "points must be (n,4) float32 or (n,3) float64");
}

// One frame: bbox -> voxel downsample (+ density grid) into the caller's preallocated outputs (FramePipeline's buffers).
// origin / xy_range: empty tensors = derived from the cloud (CPU float64 tensors of 3 / 4 values otherwise).
int64_t frame_voxel_density(const at::Tensor& points, double voxel_size, double grid_size, const at::Tensor& origin,
                         const at::Tensor& xy_range, at::Tensor voxel_key, at::Tensor inverse, at::Tensor voxels,
                         const c10::optional<at::Tensor>& grid, at::Tensor desc, at::Tensor ws, int64_t max_points,
                         int64_t max_key_space, int64_t max_nx, int64_t max_ny) {
    TORCH_CHECK(point_format(points) == LIDAR_FMT_F32X4, "frames are (n,4) float32");
    need_cuda(voxel_key, "voxel_key"); need_cuda(inverse, "inverse"); need_cuda(voxels, "voxels");
    need_cuda(desc, "desc"); need_cuda(ws, "ws");
    const c10::cuda::CUDAGuard guard(points.device());
    lidar_frame_caps caps{max_points, max_key_space, (int32_t)max_nx, (int32_t)max_ny};
    double o3[3], r4[4];
    const double* po = nullptr;
    const double* pr = nullptr;
    if (origin.numel() == 3) { auto a = origin.to(at::kCPU, at::kDouble).contiguous(); for (int i = 0; i < 3; ++i) o3[i] = a.data_ptr<double>()[i]; po = o3; }
    if (xy_range.numel() == 4) { auto a = xy_range.to(at::kCPU, at::kDouble).contiguous(); for (int i = 0; i < 4; ++i) r4[i] = a.data_ptr<double>()[i]; pr = r4; }
    int32_t* g = grid.has_value() && grid->defined() && grid->numel() ? grid->data_ptr<int32_t>() : nullptr;
    return lidar_frame_voxel_density(points.data_ptr(), points.size(0), voxel_size, grid_size, po, pr, voxel_key.data_ptr<int32_t>(),
                                     inverse.data_ptr<int32_t>(), reinterpret_cast<lidar_voxel*>(voxels.data_ptr()), g,
                                     reinterpret_cast<lidar_frame_desc*>(desc.data_ptr()), &caps, ws.data_ptr(), (size_t)ws.numel(),
                                     stream_of(points));
}

// np.histogram2d semantics on the x / y of a cloud; counts are ADDED into `counts` (int32 [nx][ny]).
int64_t hist2d_points(const at::Tensor& points, const at::Tensor& x_edges, const at::Tensor& y_edges, at::Tensor counts, int64_t mode) {
    const int fmt = point_format(points);
    need_cuda(x_edges, "x_edges"); need_cuda(y_edges, "y_edges"); need_cuda(counts, "counts");
    TORCH_CHECK(x_edges.scalar_type() == at::kDouble && y_edges.scalar_type() == at::kDouble && counts.scalar_type() == at::kInt,
                "edges are float64, counts int32");
    const c10::cuda::CUDAGuard guard(points.device());
    return lidar_hist2d_points(points.data_ptr(), fmt, points.size(0), x_edges.data_ptr<double>(), (int)x_edges.numel() - 1,
                               y_edges.data_ptr<double>(), (int)y_edges.numel() - 1, counts.data_ptr<int32_t>(), (int)mode,
                               stream_of(points));
}

// out8 <- {min x,y,z,w, max x,y,z,w} as float64; `ws` >= lidar_reduce_workspace_bytes().
int64_t bbox(const at::Tensor& points, at::Tensor out8, at::Tensor ws) {
    const int fmt = point_format(points);
    need_cuda(ws, "ws"); need_cuda(out8, "out8");
    TORCH_CHECK(out8.scalar_type() == at::kDouble && out8.numel() == 8, "out8 must hold 8 float64 values");
    const c10::cuda::CUDAGuard guard(points.device());
    return lidar_bbox(points.data_ptr(), fmt, points.size(0), out8.data_ptr<double>(), ws.data_ptr(), (size_t)ws.numel(), stream_of(points));
}

int64_t abi_version() { return lidar_abi_version(); }

}  // namespace

TORCH_LIBRARY(lidar_b200, m) {
    m.def("frame_voxel_density(Tensor points, float voxel_size, float grid_size, Tensor origin, Tensor xy_range, Tensor(a!) voxel_key, "
          "Tensor(b!) inverse, Tensor(c!) voxels, Tensor(d!)? grid, Tensor(e!) desc, Tensor(f!) ws, int max_points, int max_key_space, "
          "int max_nx, int max_ny) -> int");
    m.def("hist2d_points(Tensor points, Tensor x_edges, Tensor y_edges, Tensor(a!) counts, int mode) -> int");
    m.def("bbox(Tensor points, Tensor(a!) out8, Tensor(b!) ws) -> int");
    m.def("abi_version() -> int");
}
TORCH_LIBRARY_IMPL(lidar_b200, CUDA, m) {
    m.impl("frame_voxel_density", &frame_voxel_density);
    m.impl("hist2d_points", &hist2d_points);
    m.impl("bbox", &bbox);
}
TORCH_LIBRARY_IMPL(lidar_b200, CompositeExplicitAutograd, m) { m.impl("abi_version", &abi_version); }
