// S1a — fused bin-and-sort of points by BEV cell: one launch, one thread-block cluster per frame.
//
// There is no binning code in the reference; the cell convention is the one its target generator
// uses for box centres (src/centernet_target.py:222-224,250-257,285): voxel=(max-min)/W,
// px=(x-x_min)/voxel, reject px<0 or px>=W, ix=int(px), flat=iy*W+ix.  Here it is evaluated in fp32
// with IEEE subtract/divide (no reciprocal, no FMA) so that a numpy float32 restatement is bit-exact.
//
//   grid (CL, B), cluster (CL,1,1), 256 threads.  CTA `rank` of frame b owns the contiguous slice
//   [rank*slice, (rank+1)*slice) of the frame's points.
//     phase 1  stream the slice (x,y only), write cell[], count into a per-CTA shared-memory
//              histogram over the H*W cells (+1 overflow bin for out-of-grid points);
//     phase 2  cluster-wide exclusive scan over (cell, rank) through distributed shared memory:
//              the cells are dealt out to the CTAs, each CTA reads the CL counts of its cells from
//              its peers' histograms, scans, writes offsets[] and overwrites the peers' histogram
//              entries with the start position of (cell, rank);
//     phase 3  one warp per CTA walks its slice in point order, 32 points a step: match.any groups
//              the lanes that hit the same cell, the group takes `count` consecutive slots from the
//              cell's cursor in shared memory — a stable counting sort, no global atomics, no
//              second pass over the points, deterministic output.
//
// bin_sort_ranked_kernel is the fast form of the same sort (bin_sort_kernel stays as the any-size
// fallback).  The serial phase-3 walk is gone: in phase 1 every warp walks its OWN contiguous
// sub-slice in point order and, with the same match.any grouping, gives each point its rank among
// the points of (cell, CTA, warp) while counting into a per-warp 16-bit histogram; the scan of phase 2
// then runs over (cell, CTA) with the eight per-warp counts of a cell fetched as ONE 16-byte
// distributed-shared-memory read; and the placement is a fully parallel pass:
//     perm[ base(cell, CTA) + sum of counts(cell, CTA, warps before mine) + rank ] = point.
#include <cooperative_groups.h>

#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "lidar_prepare.cuh"

namespace b200bev {
namespace {

constexpr int kBinThreads = 256;
static_assert(kBinThreads == kPrepThreads, "the fused kernel runs lidar_prepare_frame with bin_sort's block shape");
constexpr int kMaxCluster = 8;

struct BinArgs {
  const float* pts;
  int B, N, C;
  float x_min, y_min, vx, vy;
  int W, H;
  int32_t* cell;
  int32_t* perm;
  int32_t* offsets;
  int slice;        // points per CTA, multiple of 32
  int cache_cells;  // keep the slice's cell ids in shared memory between phase 1 and 3
};

__device__ __forceinline__ int cell_of(float x, float y, const BinArgs& a) {
  const float px = __fdiv_rn(__fsub_rn(x, a.x_min), a.vx);
  const float py = __fdiv_rn(__fsub_rn(y, a.y_min), a.vy);
  // written so that NaN is rejected too
  if (!(px >= 0.0f) || !(px < (float)a.W) || !(py >= 0.0f) || !(py < (float)a.H)) return -1;
  return (int)py * a.W + (int)px;  // truncation == floor for non-negative values
}

__global__ void __launch_bounds__(kBinThreads) bin_sort_kernel(BinArgs a) {
  cg::cluster_group cluster = cg::this_cluster();
  const int CL = (int)cluster.num_blocks();
  const int rank = (int)cluster.block_rank();
  const int b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int nwarps = kBinThreads / 32;
  const int HW = a.W * a.H;
  const int nb = HW + 1;

  extern __shared__ __align__(16) uint32_t smem[];
  uint32_t* hist = smem;          // nb: count, then cursor, of (cell, this rank)
  uint32_t* cta_tot = hist + nb;  // kMaxCluster: number of points in each rank's share of the cells
  uint32_t* wsum = cta_tot + kMaxCluster;  // nwarps
  int32_t* cells_s = reinterpret_cast<int32_t*>(wsum + nwarps);

  for (int i = tid; i < nb; i += kBinThreads) hist[i] = 0;
  __syncthreads();

  const int start = rank * a.slice;
  const int end = min(a.N, start + a.slice);
  const float* pts = a.pts + (size_t)b * a.N * a.C;
  int32_t* cell_out = a.cell + (size_t)b * a.N;

  // ---- phase 1: cell ids + per-CTA histogram ----
  const bool vec2 = ((a.C & 1) == 0) && ((reinterpret_cast<uintptr_t>(a.pts) & 7) == 0);
  for (int i = start + tid; i < end; i += kBinThreads) {
    float x, y;
    if (vec2) {
      const float2 xy = __ldg(reinterpret_cast<const float2*>(pts + (size_t)i * a.C));
      x = xy.x;
      y = xy.y;
    } else {
      x = __ldg(pts + (size_t)i * a.C);
      y = __ldg(pts + (size_t)i * a.C + 1);
    }
    const int c = cell_of(x, y, a);
    cell_out[i] = c;
    if (a.cache_cells) cells_s[i - start] = c;
    atomicAdd(&hist[c < 0 ? HW : c], 1u);
  }
  cluster.sync();

  // ---- phase 2: exclusive scan over (cell, rank), cells dealt out to the CTAs ----
  const int share = ceil_div(nb, CL);
  const int lo = rank * share, hi = min(nb, lo + share);
  {
    uint32_t mine = 0;
    for (int bin = lo + tid; bin < hi; bin += kBinThreads)
      for (int q = 0; q < CL; ++q) mine += cluster.map_shared_rank(hist, q)[bin];
    for (int d = 16; d > 0; d >>= 1) mine += __shfl_xor_sync(FULL_MASK, mine, d);
    if (lane == 0) wsum[warp] = mine;
    __syncthreads();
    if (tid == 0) {
      uint32_t t = 0;
      for (int w = 0; w < nwarps; ++w) t += wsum[w];
      for (int q = 0; q < CL; ++q) cluster.map_shared_rank(cta_tot, q)[rank] = t;
    }
  }
  cluster.sync();
  uint32_t carry = 0;
  for (int q = 0; q < rank; ++q) carry += cta_tot[q];
  int32_t* offsets = a.offsets + (size_t)b * nb;
  for (int base = lo; base < hi; base += kBinThreads) {
    const int bin = base + tid;
    const bool ok = bin < hi;
    uint32_t cnt[kMaxCluster];
    uint32_t tot = 0;
#pragma unroll
    for (int q = 0; q < kMaxCluster; ++q) {
      cnt[q] = (ok && q < CL) ? cluster.map_shared_rank(hist, q)[bin] : 0u;
      tot += cnt[q];
    }
    const uint32_t incl = (uint32_t)warp_incl_scan((int)tot, lane);
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    uint32_t before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < nwarps; ++w) {
      const uint32_t v = wsum[w];
      if (w < warp) before += v;
      total += v;
    }
    if (ok) {
      uint32_t run = carry + before + incl - tot;
      offsets[bin] = (int32_t)run;
#pragma unroll
      for (int q = 0; q < kMaxCluster; ++q) {
        if (q < CL) {
          cluster.map_shared_rank(hist, q)[bin] = run;
          run += cnt[q];
        }
      }
    }
    carry += total;
    __syncthreads();
  }
  cluster.sync();

  // ---- phase 3: stable placement, one warp, slice walked in point order ----
  if (warp == 0) {
    int32_t* perm = a.perm + (size_t)b * a.N;
    const int32_t* src = a.cache_cells ? (cells_s - start) : cell_out;
    constexpr int U = 4;
    for (int i0 = start; i0 < end; i0 += 32 * U) {
      int c[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = i0 + u * 32 + lane;
        c[u] = (i < end) ? src[i] : -2;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = i0 + u * 32 + lane;
        const bool valid = i < end;
        const int bin = c[u] < 0 ? HW : c[u];
        const unsigned peers = __match_any_sync(FULL_MASK, valid ? bin : (0x40000000 | lane));
        const int rk = __popc(peers & lanemask_lt());
        const uint32_t cur = valid ? hist[bin] : 0u;
        __syncwarp();
        if (valid && rk == 0) hist[bin] = cur + (uint32_t)__popc(peers);
        __syncwarp();
        if (valid) perm[cur + rk] = i;
      }
    }
  }
  // peers may still be reading this CTA's shared memory in phase 2 only; phase 3 is local, but a
  // CTA must not exit while a peer could still address its shared memory.
  cluster.sync();
}

// ---------------------------------------------------------------------------------------------
// ranked form: all warps rank their own sub-slice in phase 1, placement is parallel
// ---------------------------------------------------------------------------------------------
struct RankedArgs {
  BinArgs a;
  int NW;   // warps (sub-slices) per CTA that walk points: 8, 4, 2 or 1
  int sub;  // points per sub-slice, multiple of 32, <= 65535
  unsigned long long* dbg;  // debug only (B200BEV_BINSORT_TRACE): clock stamps of CTA (0,0), thread 0
};

// sum of the 16-bit counts [0, upto) of one cell row (NW entries, 2*NW bytes, naturally aligned)
__device__ __forceinline__ uint32_t row_prefix(const uint16_t* row, int NW, int upto) {
  uint32_t s = 0;
  if (NW == 8) {
    const uint4 v = *reinterpret_cast<const uint4*>(row);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (2 * k < upto) s += w[k] & 0xffffu;
      if (2 * k + 1 < upto) s += w[k] >> 16;
    }
  } else {
    for (int k = 0; k < upto; ++k) s += row[k];
  }
  return s;
}

// FUSED (b200bev_lidar_prepare_bin_sort, SURVEY 8f N3 "filter + compact fused into bin-and-sort"): the frame's cluster first
// runs the range filter + stable compaction + zero padding of the raw sweep (lidar_prepare_frame writes a.pts), then — after a
// fence and a cluster barrier — bins and sorts what it wrote.  The rows were written by this launch, so phase 1 reads them
// through L2 (ld.global.cg), not through the non-coherent path.
template <bool FUSED>
__global__ void __launch_bounds__(kBinThreads) bin_sort_ranked_kernel(RankedArgs ra, PrepArgs pa) {
  const BinArgs& a = ra.a;
  int dbg_n = 0;
  auto stamp = [&]() { if (ra.dbg && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) ra.dbg[dbg_n++] = clock64(); };
  stamp();
  cg::cluster_group cluster = cg::this_cluster();
  if constexpr (FUSED) {
    __shared__ uint32_t prep_warp_cnt[kPrepThreads / 32];
    __shared__ uint32_t prep_cta_tot[kPrepMaxCluster];
    lidar_prepare_frame(pa, cluster, prep_warp_cnt, prep_cta_tot);
    __threadfence();
    cluster.sync();
  }
  const int CL = (int)cluster.num_blocks();
  const int rank = (int)cluster.block_rank();
  const int b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int nwarps = kBinThreads / 32;
  const int HW = a.W * a.H;
  const int nb = HW + 1;
  const int NW = ra.NW;

  extern __shared__ __align__(16) uint32_t smem[];
  uint32_t* base = smem;                    // nb: start of (cell, this CTA) in perm
  uint32_t* cta_tot = base + nb;            // kMaxCluster
  uint32_t* wsum = cta_tot + kMaxCluster;   // nwarps
  uint16_t* whist = reinterpret_cast<uint16_t*>(smem + (((size_t)nb + kMaxCluster + nwarps + 3) & ~(size_t)3));  // [nb][NW]
  uint16_t* rank_s = whist + (((size_t)nb * NW + 7) & ~(size_t)7);   // [slice]
  int32_t* cells_s = reinterpret_cast<int32_t*>(rank_s + (((size_t)a.slice + 1) & ~(size_t)1));  // [slice] if cached

  {
    uint32_t* z = reinterpret_cast<uint32_t*>(whist);
    const int nz = (nb * NW + 1) / 2;
    for (int i = tid; i < nz; i += kBinThreads) z[i] = 0;
  }
  __syncthreads();
  stamp();

  const int start = rank * a.slice;
  const int end = min(a.N, start + a.slice);
  const float* pts = a.pts + (size_t)b * a.N * a.C;
  int32_t* cell_out = a.cell + (size_t)b * a.N;

  // ---- phase 1: cell ids, rank within (cell, warp), per-warp histogram; sub-slice walked in order ----
  if (warp < NW) {
    const bool vec2 = ((a.C & 1) == 0) && ((reinterpret_cast<uintptr_t>(a.pts) & 7) == 0);
    const int s0 = start + warp * ra.sub;
    const int s1 = min(end, s0 + ra.sub);
    uint16_t* wh = whist + warp;
    constexpr int U = 4;
    for (int i0 = s0; i0 < s1; i0 += 32 * U) {
      float x[U], y[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = i0 + u * 32 + lane;
        x[u] = 0.0f;
        y[u] = 0.0f;
        if (i < s1) {
          if (vec2) {
            const float2* q = reinterpret_cast<const float2*>(pts + (size_t)i * a.C);
            const float2 xy = FUSED ? __ldcg(q) : __ldg(q);
            x[u] = xy.x;
            y[u] = xy.y;
          } else {
            x[u] = FUSED ? __ldcg(pts + (size_t)i * a.C) : __ldg(pts + (size_t)i * a.C);
            y[u] = FUSED ? __ldcg(pts + (size_t)i * a.C + 1) : __ldg(pts + (size_t)i * a.C + 1);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = i0 + u * 32 + lane;
        const bool valid = i < s1;
        const int c = cell_of(x[u], y[u], a);
        const int bin = c < 0 ? HW : c;
        const unsigned peers = __match_any_sync(FULL_MASK, valid ? bin : (0x40000000 | lane));
        const int rk = __popc(peers & lanemask_lt());
        const uint32_t cur = valid ? wh[(size_t)bin * NW] : 0u;
        __syncwarp();
        if (valid && rk == 0) wh[(size_t)bin * NW] = (uint16_t)(cur + (uint32_t)__popc(peers));
        __syncwarp();
        if (valid) {
          const uint32_t rnk = cur + (uint32_t)rk;
          if (a.cache_cells) {
            cell_out[i] = c;
            rank_s[i - start] = (uint16_t)rnk;
            cells_s[i - start] = c;
          } else {
            // no room in shared memory (big grids / long sweeps): the rank rides in the upper half of the cell word
            // until phase 3 (cell ids fit 16 bits: W*H <= 48000; -1 <-> 0xffff)
            cell_out[i] = (int32_t)(((uint32_t)c & 0xffffu) | (rnk << 16));
          }
        }
      }
    }
  }
  stamp();
  cluster.sync();
  stamp();

  // ---- phase 2: exclusive scan over (cell, CTA), cells dealt out to the CTAs ----
  auto count_of = [&](int q, int bin) -> uint32_t {
    const uint16_t* row = cluster.map_shared_rank(whist, q) + (size_t)bin * NW;
    return row_prefix(row, NW, NW);
  };
  const int share = ceil_div(nb, CL);
  const int lo = rank * share, hi = min(nb, lo + share);
  uint32_t cnt_keep[kMaxCluster];   // counts of this thread's first bin (the common case: share <= 256)
  {
    uint32_t mine = 0;
    for (int bin = lo + tid; bin < hi; bin += kBinThreads) {
      const bool first = bin < lo + kBinThreads;
#pragma unroll
      for (int q = 0; q < kMaxCluster; ++q) {
        const uint32_t c = q < CL ? count_of(q, bin) : 0u;
        if (first) cnt_keep[q] = c;
        mine += c;
      }
    }
    for (int d = 16; d > 0; d >>= 1) mine += __shfl_xor_sync(FULL_MASK, mine, d);
    if (lane == 0) wsum[warp] = mine;
    __syncthreads();
    if (tid == 0) {
      uint32_t t = 0;
      for (int w = 0; w < nwarps; ++w) t += wsum[w];
      for (int q = 0; q < CL; ++q) cluster.map_shared_rank(cta_tot, q)[rank] = t;
    }
  }
  stamp();
  cluster.sync();
  stamp();
  uint32_t carry = 0;
  for (int q = 0; q < rank; ++q) carry += cta_tot[q];
  int32_t* offsets = a.offsets + (size_t)b * nb;
  for (int bs = lo; bs < hi; bs += kBinThreads) {
    const int bin = bs + tid;
    const bool ok = bin < hi;
    uint32_t cnt[kMaxCluster];
    uint32_t tot = 0;
#pragma unroll
    for (int q = 0; q < kMaxCluster; ++q) {
      cnt[q] = (bs == lo) ? (ok ? cnt_keep[q] : 0u) : ((ok && q < CL) ? count_of(q, bin) : 0u);
      tot += cnt[q];
    }
    const uint32_t incl = (uint32_t)warp_incl_scan((int)tot, lane);
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    uint32_t before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < nwarps; ++w) {
      const uint32_t v = wsum[w];
      if (w < warp) before += v;
      total += v;
    }
    if (ok) {
      uint32_t run = carry + before + incl - tot;
      offsets[bin] = (int32_t)run;
#pragma unroll
      for (int q = 0; q < kMaxCluster; ++q) {
        if (q < CL) {
          cluster.map_shared_rank(base, q)[bin] = run;
          run += cnt[q];
        }
      }
    }
    carry += total;
    __syncthreads();
  }
  stamp();
  cluster.sync();   // every base[] is written; no remote access after this point
  stamp();

  // ---- phase 3: parallel placement ----
  int32_t* perm = a.perm + (size_t)b * a.N;
  for (int i = start + tid; i < end; i += kBinThreads) {
    const int li = i - start;
    int c;
    uint32_t rnk;
    if (a.cache_cells) {
      c = cells_s[li];
      rnk = rank_s[li];
    } else {
      const uint32_t packed = (uint32_t)cell_out[i];
      c = (int)(int16_t)(packed & 0xffffu);
      rnk = packed >> 16;
      cell_out[i] = c;
    }
    const int bin = c < 0 ? HW : c;
    const int w = li / ra.sub;
    const uint32_t pos = base[bin] + row_prefix(whist + (size_t)bin * NW, NW, w) + rnk;
    perm[pos] = i;
  }
  stamp();
}

}  // namespace
}  // namespace b200bev

using namespace b200bev;

namespace {
// prep == nullptr: bin-and-sort of `points`.  prep != nullptr: the raw sweeps are filtered / compacted / padded into `points`
// first — in the same launch when the ranked kernel takes the shape, as a launch of its own in front of the fallback kernel.
int launch_bin_sort(const float* points, int B, int N, int C, float x_min, float y_min, float voxel_x, float voxel_y, int W, int H,
                    int32_t* cell, int32_t* perm, int32_t* offsets, const PrepArgs* prep, int64_t max_frame_rows, const float* pc_range,
                    void* stream) {
  BinArgs a{};
  a.pts = points; a.B = B; a.N = N; a.C = C;
  a.x_min = x_min; a.y_min = y_min; a.vx = voxel_x; a.vy = voxel_y; a.W = W; a.H = H;
  a.cell = cell; a.perm = perm; a.offsets = offsets;

  int CL = 1;
  while (CL < kMaxCluster && N / (CL * 2) >= 1024) CL *= 2;
  a.slice = ceil_div(ceil_div(N, CL), 32) * 32;
  const size_t nb = (size_t)W * H + 1;

  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(CL, B, 1);
  cfg.blockDim = dim3(kBinThreads, 1, 1);
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;

  // ranked form: the widest per-warp histogram that fits shared memory (and 16-bit counts)
  const char* force = debug_env("B200BEV_BINSORT");
  const bool legacy = force && force[0] == 'l';
  if (!legacy) {
    const size_t kLimit = 220 * 1024;
    for (int NW = kBinThreads / 32; NW >= 1; NW >>= 1) {
      const int sub = ceil_div(ceil_div(a.slice, NW), 32) * 32;
      if (sub > 65535) break;   // halving NW only makes it larger
      const size_t head = ((nb + kMaxCluster + kBinThreads / 32 + 3) & ~(size_t)3) * sizeof(uint32_t);
      const size_t hist = ((nb * NW + 7) & ~(size_t)7) * sizeof(uint16_t);
      const size_t cache = (((size_t)a.slice + 1) & ~(size_t)1) * sizeof(uint16_t) + (size_t)a.slice * sizeof(int32_t);
      size_t smem = head + hist;
      if (smem > kLimit) continue;
      // ranks and cell ids of the slice stay in shared memory when two CTAs per SM still fit; otherwise they
      // travel packed in the `cell` output (16 + 16 bits) and all of shared memory goes to the per-warp histograms
      a.cache_cells = (smem + cache <= 100 * 1024 && (long long)W * H < 0xffff) ? 1 : 0;
      if (a.cache_cells) smem += cache;
      else if ((long long)W * H >= 0xffff) break;   // cell ids would not fit the packed form: legacy kernel
      RankedArgs ra{a, NW, sub, nullptr};
      if (debug_env("B200BEV_BINSORT_TRACE")) B200BEV_CUDA_TRY(cudaMalloc(&ra.dbg, 16 * sizeof(unsigned long long)));
      auto kern = prep ? bin_sort_ranked_kernel<true> : bin_sort_ranked_kernel<false>;
      if (smem > 48 * 1024) B200BEV_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      cfg.dynamicSmemBytes = smem;
      B200BEV_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, ra, prep ? *prep : PrepArgs{}));
      if (ra.dbg) {   // debug: phase boundaries of CTA (0,0) in SM clocks since its start
        unsigned long long h[16];
        B200BEV_CUDA_TRY(cudaStreamSynchronize(cfg.stream));
        B200BEV_CUDA_TRY(cudaMemcpy(h, ra.dbg, sizeof(h), cudaMemcpyDeviceToHost));
        fprintf(stderr, "bin_sort trace (clk since start): zero %llu | phase1 %llu | sync %llu | mine %llu | sync %llu | scan %llu | sync %llu | place %llu\n",
                h[1] - h[0], h[2] - h[0], h[3] - h[0], h[4] - h[0], h[5] - h[0], h[6] - h[0], h[7] - h[0], h[8] - h[0]);
        cudaFree(ra.dbg);
      }
      return launch_status();
    }
  }

  if (prep) {   // shapes the ranked kernel does not take: the two steps as two launches
    const int rc = b200bev_lidar_prepare(prep->raw, prep->frame_off, B, C, max_frame_rows, pc_range, N, nullptr, prep->out, prep->count,
                                         nullptr, 0, stream);
    if (rc != B200BEV_OK) return rc;
  }
  const size_t fixed = (nb + kMaxCluster + kBinThreads / 32) * sizeof(uint32_t);
  size_t smem = fixed + (size_t)a.slice * sizeof(int32_t);
  a.cache_cells = 1;
  if (smem > 200 * 1024) {
    a.cache_cells = 0;
    smem = fixed;
  }
  if (smem > 48 * 1024)
    B200BEV_CUDA_TRY(cudaFuncSetAttribute(bin_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cfg.dynamicSmemBytes = smem;
  B200BEV_CUDA_TRY(cudaLaunchKernelEx(&cfg, bin_sort_kernel, a));
  return launch_status();
}
}  // namespace

extern "C" B200BEV_API int b200bev_bin_sort(const float* points, int B, int N, int C, float x_min, float y_min, float voxel_x,
                                float voxel_y, int W, int H, int32_t* cell, int32_t* perm, int32_t* offsets,
                                void* stream) {
  if (!points || !cell || !perm || !offsets || B <= 0 || N <= 0 || C < 2 || W <= 0 || H <= 0) return B200BEV_ERR_INVALID_ARGUMENT;
  if (!(voxel_x > 0.0f) || !(voxel_y > 0.0f)) return B200BEV_ERR_INVALID_ARGUMENT;
  if ((long long)W * H > 48000 || B > 65535) return B200BEV_ERR_UNSUPPORTED;
  return launch_bin_sort(points, B, N, C, x_min, y_min, voxel_x, voxel_y, W, H, cell, perm, offsets, nullptr, 0, nullptr, stream);
}

extern "C" B200BEV_API int b200bev_lidar_prepare_bin_sort(const float* raw, const int64_t* frame_offsets, int B, int C,
                                              int64_t max_frame_rows, const float* pc_range, int max_points, float voxel_x,
                                              float voxel_y, int W, int H, float* points, int32_t* count, int32_t* cell,
                                              int32_t* perm, int32_t* offsets, void* stream) {
  if (!raw || !frame_offsets || !pc_range || !points || !count || !cell || !perm || !offsets || B <= 0 || C < 3 || max_points <= 0 ||
      max_frame_rows < 0 || W <= 0 || H <= 0)
    return B200BEV_ERR_INVALID_ARGUMENT;
  if (!(voxel_x > 0.0f) || !(voxel_y > 0.0f)) return B200BEV_ERR_INVALID_ARGUMENT;
  if ((long long)W * H > 48000 || B > 65535 || max_frame_rows > 0x7fffffffLL) return B200BEV_ERR_UNSUPPORTED;
  PrepArgs pa{};
  pa.raw = raw; pa.frame_off = frame_offsets; pa.B = B; pa.C = C; pa.max_points = max_points;
  for (int i = 0; i < 3; ++i) {
    pa.lo[i] = pc_range[i];
    pa.hi[i] = pc_range[3 + i];
  }
  pa.select = nullptr; pa.out = points; pa.count = count; pa.kept_index = nullptr; pa.cap = max_frame_rows;
  return launch_bin_sort(points, B, max_points, C, pc_range[0], pc_range[1], voxel_x, voxel_y, W, H, cell, perm, offsets, &pa,
                         max_frame_rows, pc_range, stream);
}
