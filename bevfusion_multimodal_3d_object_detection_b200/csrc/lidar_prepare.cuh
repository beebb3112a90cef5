// N3 — range filter + pad / subsample of raw LiDAR sweeps: argument block and the per-frame device body, shared by
// lidar_prepare.cu (its own launch) and bin_sort.cu (the same work in front of the bin-and-sort phases, one launch).
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace b200bev {

constexpr int kPrepThreads = 256;   // == kBinThreads of bin_sort.cu
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

// One frame (blockIdx.y) by one cluster of 256-thread CTAs.  A __device__ function so that bin_sort.cu can run it in front
// of its own phases in the same launch (b200bev_lidar_prepare_bin_sort).  warp_cnt: [8], cta_tot: [8] words of shared memory.
__device__ __forceinline__ void lidar_prepare_frame(const PrepArgs& a, cg::cluster_group& cluster, uint32_t* warp_cnt, uint32_t* cta_tot) {
  const int CL = (int)cluster.num_blocks();
  const int rank = (int)cluster.block_rank();
  const int b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int nwarps = kPrepThreads / 32;

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


}  // namespace b200bev
