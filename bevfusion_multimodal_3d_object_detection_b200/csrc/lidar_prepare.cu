// N3 — the step in front of S1: range filter + pad / subsample of raw LiDAR sweeps.
//
// Reference: NuScenesDataset._load_lidar_points and _pad_or_subsample, src/train_detect.py:147-189 —
//   mask = (x > x_min) & (x < x_max) & (y > y_min) & (y < y_max) & (z > z_min) & (z < z_max)   (strict; NaN fails)
//   points = points[mask]                                  (file order kept)
//   N <  max_points: append zero rows up to max_points
//   N >= max_points: points[np.random.choice(N, max_points, replace=False)]
// done per sample in numpy on a DataLoader worker.  Here a whole batch of raw sweeps (rows of all frames
// back to back, CSR frame offsets) is compacted in one launch:
//
//   grid (CL, B), one thread-block cluster per frame, 256 threads; CTA `rank` owns a contiguous slice.
//     pass 1  every warp counts the in-range points of its contiguous share (ballot + popc);
//     scan    warp counts -> CTA total -> cluster-wide exclusive scan through distributed shared memory;
//     pass 2  the slice is read again (it is L2-resident by now), in-range points are written at
//             base + rank-in-warp: a stable compaction, so the output order is the file order;
//     pad     the cluster zero-fills rows [kept, max_points).
//   `select` (optional, (B,max_points) i32) reproduces the subsample branch: row j of the output is the
//   select[b,j]-th in-range point; the caller draws the indices (np.random.choice) — the reference's draw
//   is unseeded, so no implementation can reproduce it without being handed the indices.  Without
//   `select`, a frame with more than max_points in-range points keeps the first max_points.
//   count[b] always reports the number of in-range points of the frame.
#include "lidar_prepare.cuh"

namespace b200bev {
namespace {

__global__ void __launch_bounds__(kPrepThreads) lidar_prepare_kernel(PrepArgs a) {
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ uint32_t warp_cnt[kPrepThreads / 32];
  __shared__ uint32_t cta_tot[kPrepMaxCluster];
  lidar_prepare_frame(a, cluster, warp_cnt, cta_tot);
}

}  // namespace
}  // namespace b200bev

using namespace b200bev;

extern "C" B200BEV_API size_t b200bev_lidar_prepare_workspace_bytes(int B, int64_t max_frame_rows, int with_select) {
  if (B <= 0 || max_frame_rows < 0 || !with_select) return 0;
  return (size_t)B * (size_t)max_frame_rows * sizeof(int32_t);
}

extern "C" B200BEV_API int b200bev_lidar_prepare(const float* raw, const int64_t* frame_offsets, int B, int C,
                                                 int64_t max_frame_rows, const float* pc_range, int max_points,
                                                 const int32_t* select, float* out, int32_t* count, void* workspace,
                                                 size_t workspace_bytes, void* stream) {
  if (!raw || !frame_offsets || !pc_range || !out || !count || B <= 0 || C < 3 || max_points <= 0 || max_frame_rows < 0)
    return B200BEV_ERR_INVALID_ARGUMENT;
  if (B > 65535 || max_frame_rows > 0x7fffffffLL) return B200BEV_ERR_UNSUPPORTED;
  if (select && (!workspace || workspace_bytes < b200bev_lidar_prepare_workspace_bytes(B, max_frame_rows, 1)))
    return B200BEV_ERR_WORKSPACE;
  PrepArgs a{};
  a.raw = raw; a.frame_off = frame_offsets; a.B = B; a.C = C; a.max_points = max_points;
  for (int i = 0; i < 3; ++i) {
    a.lo[i] = pc_range[i];
    a.hi[i] = pc_range[3 + i];
  }
  a.select = select; a.out = out; a.count = count;
  a.kept_index = reinterpret_cast<int32_t*>(workspace);
  a.cap = max_frame_rows;
  int CL = 1;
  while (CL < kPrepMaxCluster && max_frame_rows / (CL * 2) >= 2048) CL *= 2;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(CL, B, 1);
  cfg.blockDim = dim3(kPrepThreads, 1, 1);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  B200BEV_CUDA_TRY(cudaLaunchKernelEx(&cfg, lidar_prepare_kernel, a));
  return launch_status();
}
