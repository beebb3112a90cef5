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
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace b200bev {
namespace {

constexpr int kPrepThreads = 256;
constexpr int kPrepMaxCluster = 8;

struct PrepArgs {
  const float* raw;          // (total, C)
  const int64_t* frame_off;  // (B+1)
  int B, C, max_points;
  float lo[3], hi[3];
  const int32_t* select;     // (B, max_points) or nullptr
  float* out;                // (B, max_points, C)
  int32_t* count;            // (B)
  int32_t* kept_index;       // workspace (B, cap) — only with `select`: position -> raw row of the k-th in-range point
  long long cap;             // rows of kept_index per frame
};

__device__ __forceinline__ bool in_range(float x, float y, float z, const PrepArgs& a) {
  return (x > a.lo[0]) && (x < a.hi[0]) && (y > a.lo[1]) && (y < a.hi[1]) && (z > a.lo[2]) && (z < a.hi[2]);
}

__global__ void __launch_bounds__(kPrepThreads) lidar_prepare_kernel(PrepArgs a) {
  cg::cluster_group cluster = cg::this_cluster();
  const int CL = (int)cluster.num_blocks();
  const int rank = (int)cluster.block_rank();
  const int b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int nwarps = kPrepThreads / 32;

  __shared__ uint32_t warp_cnt[nwarps];
  __shared__ uint32_t cta_tot[kPrepMaxCluster];

  const long long f0 = a.frame_off[b], f1 = a.frame_off[b + 1];
  const long long n = f1 - f0;
  // slices and warp shares are multiples of 32 rows, so a warp always reads 32 consecutive rows
  const long long slice = ((n + CL - 1) / CL + 31) / 32 * 32;
  const long long s_begin = min(n, (long long)rank * slice), s_end = min(n, s_begin + slice);
  const long long share = ((s_end - s_begin + nwarps - 1) / nwarps + 31) / 32 * 32;
  const long long w_begin = min(s_end, s_begin + (long long)warp * share), w_end = min(s_end, w_begin + share);
  const float* raw = a.raw + f0 * a.C;
  const bool vec4 = a.C == 4 && (reinterpret_cast<uintptr_t>(a.raw) & 15) == 0;

  auto keep_row = [&](long long i) -> bool {
    if (vec4) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(raw + i * 4));
      return in_range(v.x, v.y, v.z, a);
    }
    const float* p = raw + i * a.C;
    return in_range(__ldg(p), __ldg(p + 1), __ldg(p + 2), a);
  };

  // ---- pass 1: count ----
  uint32_t mine = 0;
  for (long long i0 = w_begin; i0 < w_end; i0 += 32) {
    const long long i = i0 + lane;
    const bool k = i < w_end && keep_row(i);
    mine += __popc(__ballot_sync(FULL_MASK, k));
  }
  if (lane == 0) warp_cnt[warp] = mine;   // every lane holds the same total
  __syncthreads();
  uint32_t before_warp = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < nwarps; ++w) {
    const uint32_t v = warp_cnt[w];
    if (w < warp) before_warp += v;
    tot += v;
  }
  if (tid == 0)
    for (int q = 0; q < CL; ++q) cluster.map_shared_rank(cta_tot, q)[rank] = tot;
  cluster.sync();
  uint32_t base = before_warp, kept = 0;
  for (int q = 0; q < CL; ++q) {
    const uint32_t v = cta_tot[q];
    if (q < rank) base += v;
    kept += v;
  }
  if (rank == 0 && tid == 0) a.count[b] = (int32_t)kept;

  float* out = a.out + (size_t)b * a.max_points * a.C;
  const bool gather = a.select != nullptr;
  int32_t* kept_index = gather ? a.kept_index + (size_t)b * a.cap : nullptr;

  // ---- pass 2: stable placement ----
  uint32_t pos = base;
  for (long long i0 = w_begin; i0 < w_end; i0 += 32) {
    const long long i = i0 + lane;
    bool k = false;
    float4 v4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < w_end) {
      if (vec4) {
        v4 = __ldg(reinterpret_cast<const float4*>(raw + i * 4));
        k = in_range(v4.x, v4.y, v4.z, a);
      } else {
        k = keep_row(i);
      }
    }
    const unsigned bal = __ballot_sync(FULL_MASK, k);
    const uint32_t dst = pos + __popc(bal & lanemask_lt());
    if (k) {
      if (gather) {
        kept_index[dst] = (int32_t)i;
      } else if (dst < (uint32_t)a.max_points) {
        if (vec4) {
          *reinterpret_cast<float4*>(out + (size_t)dst * 4) = v4;
        } else {
          for (int c = 0; c < a.C; ++c) out[(size_t)dst * a.C + c] = __ldg(raw + i * a.C + c);
        }
      }
    }
    pos += __popc(bal);
  }

  if (!gather) {
    // ---- pad: zero rows [kept, max_points), dealt out over the cluster ----
    const long long z0 = (long long)min(kept, (uint32_t)a.max_points) * a.C, z1 = (long long)a.max_points * a.C;
    for (long long j = z0 + (long long)rank * kPrepThreads + tid; j < z1; j += (long long)CL * kPrepThreads) out[j] = 0.0f;
    return;
  }

  // ---- subsample branch: out[j] = the select[j]-th in-range point (all of kept_index must be written first) ----
  __threadfence();
  cluster.sync();
  const int32_t* sel = a.select + (size_t)b * a.max_points;
  for (long long j = (long long)rank * kPrepThreads + tid; j < a.max_points; j += (long long)CL * kPrepThreads) {
    const int32_t s = __ldg(sel + j);
    if (s >= 0 && (uint32_t)s < kept) {
      const long long i = __ldcg(kept_index + s);
      for (int c = 0; c < a.C; ++c) out[(size_t)j * a.C + c] = __ldg(raw + i * a.C + c);
    } else {
      for (int c = 0; c < a.C; ++c) out[(size_t)j * a.C + c] = 0.0f;   // index past the frame: a padding row
    }
  }
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
