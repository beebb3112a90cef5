// S2 (geometric form) — camera features -> BEV canvas by calibrated projection.
//
// north_star S2: "projects BEV cell centres through per-camera intrinsics/extrinsics and bilinearly
// gathers ResNet-18 image features".  The reference has no geometric projection; the
// arithmetic is the restatement in oracle/bev_oracle.py (project_cells / camera_project): pinhole
// projection of the cell centre on the z = z_plane ground plane, grid_sample(align_corners=False,
// padding_mode='zeros') taps, mean over the cameras that see the cell.  Calibration layout as written
// by src/data_converter.py:110-117 (per camera: 3x3 intrinsic, ego->camera [R|t]).
//
// Two kernels compute the same bits:
//
//   camera_project_staged_kernel   the fast path.  The features are NCHW, so the rows of one (camera,
//       channel) plane that any BEV cell samples form ONE contiguous byte range ("band": ground cells
//       project below the horizon, ~40 % of the plane).  A persistent CTA per SM walks (frame, channel
//       group) items; a producer thread streams the bands of the item's cameras into a shared-memory
//       ring with cp.async.bulk (TMA 1-D) + mbarrier expect_tx, 15 consumer warps gather the bilinear
//       taps out of shared memory.  Each consumer thread owns fixed cells (lane <-> consecutive cells),
//       keeps their camera-visibility masks and accumulators in registers, reads the (u,v) table the
//       CTA built once in shared memory, and writes every canvas value exactly once, coalesced.
//       HBM sees only full-line sequential reads and coalesced writes.
//   camera_project_gather_kernel   the general fallback (any shape): taps read straight from global
//       memory, lanes <-> cells.
#include <algorithm>
#include <climits>
#include <cstdlib>

#include "async_copy.cuh"
#include "common.cuh"

namespace b200bev {
namespace {

constexpr int kMaxCams = 8;
constexpr float kNearPlane = 0.1f;  // metres in front of the camera

struct ProjArgs {
  const float* feats;
  int B, n_cam, C, h, w;
  const float* K;   // (T,n_cam,3,3)
  const float* E;   // (T,n_cam,3,4)
  int T;
  float img_w, img_h, x_min, y_min, vx, vy, z_plane;
  int W, H;
  float* out;
  float* uv_valid;
};

// Feature-map coordinates (u,v) of BEV cell (ix,iy) in one camera; every operation is a single
// correctly-rounded fp32 op in a fixed order so that numpy float32 reproduces it bit for bit.
__device__ __forceinline__ bool project_cell(const ProjArgs& a, const float* __restrict__ Kc, const float* __restrict__ Ec,
                                             int ix, int iy, float& u, float& v) {
  const float X = __fadd_rn(a.x_min, __fmul_rn(__fadd_rn((float)ix, 0.5f), a.vx));
  const float Y = __fadd_rn(a.y_min, __fmul_rn(__fadd_rn((float)iy, 0.5f), a.vy));
  const float Z = a.z_plane;
  float pc[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const float s = __fadd_rn(__fadd_rn(__fmul_rn(Ec[r * 4 + 0], X), __fmul_rn(Ec[r * 4 + 1], Y)), __fmul_rn(Ec[r * 4 + 2], Z));
    pc[r] = __fadd_rn(s, Ec[r * 4 + 3]);
  }
  const bool front = pc[2] > kNearPlane;
  const float zs = front ? pc[2] : 1.0f;
  const float xn = __fdiv_rn(pc[0], zs), yn = __fdiv_rn(pc[1], zs);
  const float U = __fadd_rn(__fadd_rn(__fmul_rn(Kc[0], xn), __fmul_rn(Kc[1], yn)), Kc[2]);
  const float V = __fadd_rn(__fadd_rn(__fmul_rn(Kc[3], xn), __fmul_rn(Kc[4], yn)), Kc[5]);
  const bool inside = front && (U >= 0.0f) && (U < a.img_w) && (V >= 0.0f) && (V < a.img_h);
  // pixel -> feature coordinate, grid_sample(align_corners=False): u = U * (w / img_w) - 0.5
  u = __fsub_rn(__fmul_rn(U, __fdiv_rn((float)a.w, a.img_w)), 0.5f);
  v = __fsub_rn(__fmul_rn(V, __fdiv_rn((float)a.h, a.img_h)), 0.5f);
  return inside;
}

struct Tap {
  int o00, o01, o10, o11;   // clamped offsets inside one (h,w) plane; o00 is the smallest, o11 the largest
  float w00, w01, w10, w11; // bilinear weights, 0 for taps outside the map (padding_mode='zeros')
};

__device__ __forceinline__ Tap make_tap(float u, float v, int h, int w) {
  const float fx = floorf(u), fy = floorf(v);
  const int x0 = (int)fx, y0 = (int)fy;
  const float ax = __fsub_rn(u, fx), ay = __fsub_rn(v, fy);  // weight of the +1 tap
  const float bx = __fsub_rn(__fadd_rn(fx, 1.0f), u), by = __fsub_rn(__fadd_rn(fy, 1.0f), v);
  const bool x0ok = x0 >= 0 && x0 < w, x1ok = x0 + 1 >= 0 && x0 + 1 < w;
  const bool y0ok = y0 >= 0 && y0 < h, y1ok = y0 + 1 >= 0 && y0 + 1 < h;
  const int cx0 = min(max(x0, 0), w - 1), cx1 = min(max(x0 + 1, 0), w - 1);
  const int cy0 = min(max(y0, 0), h - 1), cy1 = min(max(y0 + 1, 0), h - 1);
  Tap tp;
  tp.o00 = cy0 * w + cx0; tp.o01 = cy0 * w + cx1;
  tp.o10 = cy1 * w + cx0; tp.o11 = cy1 * w + cx1;
  tp.w00 = (x0ok && y0ok) ? __fmul_rn(bx, by) : 0.0f;
  tp.w01 = (x1ok && y0ok) ? __fmul_rn(ax, by) : 0.0f;
  tp.w10 = (x0ok && y1ok) ? __fmul_rn(bx, ay) : 0.0f;
  tp.w11 = (x1ok && y1ok) ? __fmul_rn(ax, ay) : 0.0f;
  return tp;
}

// value of one camera at one cell: ((w00*v00 + w01*v01) + w10*v10) + w11*v11, each op rounded
__device__ __forceinline__ float tap_value(const Tap& tp, const float* __restrict__ plane) {
  float val = __fmul_rn(tp.w00, plane[tp.o00]);
  val = __fadd_rn(val, __fmul_rn(tp.w01, plane[tp.o01]));
  val = __fadd_rn(val, __fmul_rn(tp.w10, plane[tp.o10]));
  val = __fadd_rn(val, __fmul_rn(tp.w11, plane[tp.o11]));
  return val;
}

__device__ __forceinline__ uint32_t pack_offsets(const Tap& tp) {
  return (uint32_t)tp.o00 | ((uint32_t)(tp.o01 - tp.o00) << 30) | ((uint32_t)(tp.o10 != tp.o00) << 31);
}

// ---------------------------------------------------------------------------------------------
// optional table output (T, H*W, n_cam, 3): u, v, valid
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) camera_project_table_kernel(ProjArgs a) {
  const int HW = a.H * a.W;
  const int cell = blockIdx.x * blockDim.x + threadIdx.x;
  const int t = blockIdx.y;
  if (cell >= HW) return;
  const int iy = cell / a.W, ix = cell % a.W;
  for (int cam = 0; cam < a.n_cam; ++cam) {
    float u, v;
    const bool vis = project_cell(a, a.K + ((size_t)t * a.n_cam + cam) * 9, a.E + ((size_t)t * a.n_cam + cam) * 12, ix, iy, u, v);
    float* o = a.uv_valid + (((size_t)t * HW + cell) * a.n_cam + cam) * 3;
    o[0] = u; o[1] = v; o[2] = vis ? 1.0f : 0.0f;
  }
}

// ---------------------------------------------------------------------------------------------
// fallback: gather straight from global memory
// ---------------------------------------------------------------------------------------------
constexpr int kCellsPerBlock = 64;

__global__ void __launch_bounds__(256) camera_project_gather_kernel(ProjArgs a) {
  __shared__ Tap taps[kCellsPerBlock][kMaxCams];
  __shared__ int cam_of[kCellsPerBlock][kMaxCams];
  __shared__ int n_vis[kCellsPerBlock];

  const int b = blockIdx.y;
  const int cell0 = blockIdx.x * kCellsPerBlock;
  const int HW = a.H * a.W;
  const int tid = threadIdx.x;
  const int t = a.T == 1 ? 0 : b;

  // stage 1: per-cell tap table (64 cells x n_cam), visible cameras compacted in camera order
  if (tid < kCellsPerBlock) {
    const int cell = cell0 + tid;
    int nv = 0;
    if (cell < HW) {
      const int iy = cell / a.W, ix = cell % a.W;
      for (int cam = 0; cam < a.n_cam; ++cam) {
        float u, v;
        if (project_cell(a, a.K + ((size_t)t * a.n_cam + cam) * 9, a.E + ((size_t)t * a.n_cam + cam) * 12, ix, iy, u, v)) {
          taps[tid][nv] = make_tap(u, v, a.h, a.w);
          cam_of[tid][nv] = cam;
          ++nv;
        }
      }
    }
    n_vis[tid] = nv;
  }
  __syncthreads();

  // stage 2: lanes <-> cells (coalesced canvas writes), warps stride over channels
  const int lane_cell = tid & (kCellsPerBlock - 1);
  const int cgroup = tid / kCellsPerBlock;              // 0..3
  constexpr int kGroups = 256 / kCellsPerBlock;
  const int cell = cell0 + lane_cell;
  if (cell >= HW) return;
  const int nv = n_vis[lane_cell];
  const size_t plane = (size_t)a.h * a.w;
  const float* fb = a.feats + (size_t)b * a.n_cam * a.C * plane;
  float* ob = a.out + (size_t)b * a.C * HW + cell;
  const float den = (float)(nv > 0 ? nv : 1);
  for (int c = cgroup; c < a.C; c += kGroups) {
    float s = 0.0f;
    for (int k = 0; k < nv; ++k) s = __fadd_rn(s, tap_value(taps[lane_cell][k], fb + ((size_t)cam_of[lane_cell][k] * a.C + c) * plane));
    ob[(size_t)c * HW] = __fdiv_rn(s, den);
  }
}

// ---------------------------------------------------------------------------------------------
// fast path: bands staged in shared memory by the copy engine
// ---------------------------------------------------------------------------------------------
constexpr int kConsumers = 480;                 // 15 gather warps: with the producer warp 16 warps, 4 per scheduler -> 128 registers each
constexpr int kStagedThreads = kConsumers + 32; // + the producer warp
constexpr int kCPT = 6;                         // cells per consumer thread
constexpr int kPartCells = kConsumers * kCPT;   // cells one CTA owns at a time (2880); larger grids are cut into parts
constexpr int kMaxStages = 16;
constexpr int kCtrlBytes = 1024;                // barriers, band bounds, calibration
constexpr int kEntryBytes = 20;                 // table entry: four tap weights + packed offsets
constexpr int kMaxSmemOptin = 232448;           // 227 KB

struct StagedCtrl {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  int n_entries;
  int band_lo[kMaxCams];   // first / one-past-last float of the plane any visible cell of the part taps
  int band_hi[kMaxCams];
  float K[kMaxCams * 9];
  float E[kMaxCams * 12];
};
static_assert(sizeof(StagedCtrl) <= kCtrlBytes, "control block must fit its slot");

// Work item = (cell part, frame, group of CG channels); a CTA walks a contiguous range of items.
template <int CG>
__global__ void __launch_bounds__(kStagedThreads, 1) camera_project_staged_kernel(ProjArgs a, int n_parts, int part_cells, int tab_cap,
                                                                                   int ring_floats) {
  constexpr int CPT = kCPT;
  extern __shared__ __align__(128) uint8_t smem[];
  StagedCtrl* ctrl = reinterpret_cast<StagedCtrl*>(smem);
  // table of every visible (cell, camera) pair: the four bilinear weights and the packed tap offsets
  // (bits 0..29 offset of the top-left tap, bit 30: the right taps are one float on, bit 31: the lower taps one row on)
  float4* wtab = reinterpret_cast<float4*>(smem + kCtrlBytes);
  uint32_t* otab = reinterpret_cast<uint32_t*>(smem + kCtrlBytes + (size_t)tab_cap * sizeof(float4));
  float* ring = reinterpret_cast<float*>(smem + kCtrlBytes + (size_t)tab_cap * kEntryBytes);
  // scratch of the table build, aliased onto the (then idle) ring
  int* hist = reinterpret_cast<int*>(ring);                               // [256] cells per visibility mask, then cursors
  uint16_t* order = reinterpret_cast<uint16_t*>(hist + 256);              // [kPartCells] cells sorted by visibility mask
  uint8_t* mask_of = reinterpret_cast<uint8_t*>(order + kPartCells);      // [kPartCells]
  float2* uv_of = reinterpret_cast<float2*>(mask_of + kPartCells);        // [part cells][n_cam] (u,v), when it fits
  constexpr int kScratchFixed = 256 * 4 + kPartCells * 3;                 // multiple of 8

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int HW = a.H * a.W;
  const int plane = a.h * a.w;
  const int n_groups = ceil_div(a.C, CG);
  const long long per_part = (long long)a.B * n_groups;
  const long long items = per_part * n_parts;
  const long long per = ceil_div64(items, (long long)gridDim.x);
  long long it = per * blockIdx.x;
  const long long it_end = min(items, it + per);
  bool first_segment = true;

  // One pass of this loop per (cell part, calibration) in use.
  while (it < it_end) {
    const int part = (int)(it / per_part);
    const int b0 = (int)((it % per_part) / n_groups);
    const long long seg_end = a.T == 1 ? min(it_end, (long long)(part + 1) * per_part)
                                       : min(it_end, (long long)part * per_part + (long long)(b0 + 1) * n_groups);
    const int t = a.T == 1 ? 0 : b0;
    const int cell_lo = part * part_cells;
    const int n_cells = min(part_cells, HW - cell_lo);

    // ---- calibration, barriers ----
    if (tid < a.n_cam * 9) ctrl->K[tid] = __ldg(a.K + (size_t)t * a.n_cam * 9 + tid);
    if (tid >= 128 && tid < 128 + a.n_cam * 12) ctrl->E[tid - 128] = __ldg(a.E + (size_t)t * a.n_cam * 12 + tid - 128);
    if (tid >= 256) hist[tid - 256] = 0;
    if (tid == 0) {
      for (int s = 0; s < kMaxStages; ++s) {
        if (!first_segment) { mbarrier_inval(&ctrl->full[s]); mbarrier_inval(&ctrl->empty[s]); }
        mbarrier_init(&ctrl->full[s], 1);
        mbarrier_init(&ctrl->empty[s], kConsumers / 32);
      }
      ctrl->n_entries = 0;
      for (int c = 0; c < kMaxCams; ++c) { ctrl->band_lo[c] = INT_MAX; ctrl->band_hi[c] = 0; }
      mbarrier_init_fence();
    }
    first_segment = false;
    __syncthreads();

    // ---- cells sorted by visibility mask: the 32 cells of a warp then (mostly) see the same camera, so a
    //      gather step runs with all lanes on instead of a few ----
    const bool uv_cached = kScratchFixed + (long long)n_cells * a.n_cam * (long long)sizeof(float2) <= (long long)ring_floats * 4;
    if (tid < kConsumers) {
      for (int i = tid; i < n_cells; i += kConsumers) {
        const int cell = cell_lo + i, iy = cell / a.W, ix = cell % a.W;
        uint32_t vis = 0;
        for (int cam = 0; cam < a.n_cam; ++cam) {
          float u, v;
          if (project_cell(a, ctrl->K + cam * 9, ctrl->E + cam * 12, ix, iy, u, v)) vis |= 1u << cam;
          if (uv_cached) uv_of[i * a.n_cam + cam] = make_float2(u, v);
        }
        mask_of[i] = (uint8_t)vis;
        atomicAdd(&hist[vis], 1);
      }
    }
    __syncthreads();
    if (warp == 0) {   // exclusive scan of the 256 bins, 8 per lane
      int v[8], sum = 0;
#pragma unroll
      for (int k = 0; k < 8; ++k) { v[k] = hist[lane * 8 + k]; sum += v[k]; }
      int run = warp_incl_scan(sum, lane) - sum;
#pragma unroll
      for (int k = 0; k < 8; ++k) { hist[lane * 8 + k] = run; run += v[k]; }
    }
    __syncthreads();
    if (tid < kConsumers)
      for (int i = tid; i < n_cells; i += kConsumers) order[atomicAdd(&hist[mask_of[i]], 1)] = (uint16_t)i;
    __syncthreads();

    // ---- table: visibility masks, entry bases, cell ids and mean factors in registers; taps in shared memory ----
    uint32_t info[CPT];   // bits 0..7: cameras that see the cell; bits 8..31: index of the cell's first table entry
    int cell_of[CPT];     // cell owned in slot j, -1 if none
    float scale[CPT];     // 1/n for n = 0, 1 or a power of two cameras (exact); 0 = "needs the IEEE division"
    if (tid < kConsumers) {
      int lo_r[kMaxCams], hi_r[kMaxCams];
#pragma unroll
      for (int cam = 0; cam < kMaxCams; ++cam) { lo_r[cam] = INT_MAX; hi_r[cam] = 0; }
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        const int slot = tid + j * kConsumers;
        const int i = slot < n_cells ? (int)order[slot] : -1;
        const int cell = i >= 0 ? cell_lo + i : -1;
        cell_of[j] = cell;
        const uint32_t vis = i >= 0 ? (uint32_t)mask_of[i] : 0u;
        const int nv = __popc(vis);
        const int base = nv ? atomicAdd(&ctrl->n_entries, nv) : 0;
        info[j] = vis | ((uint32_t)base << 8);
        scale[j] = nv <= 1 ? 1.0f : ((nv & (nv - 1)) == 0 ? __fdiv_rn(1.0f, (float)nv) : 0.0f);
        int k = 0;
#pragma unroll
        for (int cam = 0; cam < kMaxCams; ++cam) {
          if ((vis >> cam) & 1u) {
            float2 uv;
            if (uv_cached) uv = uv_of[i * a.n_cam + cam];
            else project_cell(a, ctrl->K + cam * 9, ctrl->E + cam * 12, cell % a.W, cell / a.W, uv.x, uv.y);
            const Tap tp = make_tap(uv.x, uv.y, a.h, a.w);
            if (base + k < tab_cap) {
              wtab[base + k] = make_float4(tp.w00, tp.w01, tp.w10, tp.w11);
              otab[base + k] = pack_offsets(tp);
            }
            ++k;
            lo_r[cam] = min(lo_r[cam], tp.o00);
            hi_r[cam] = max(hi_r[cam], tp.o11 + 1);
          }
        }
      }
#pragma unroll
      for (int cam = 0; cam < kMaxCams; ++cam) {
        const int lo = __reduce_min_sync(FULL_MASK, lo_r[cam]);
        const int hi = __reduce_max_sync(FULL_MASK, hi_r[cam]);
        if (lane == 0 && hi > 0) {
          atomicMin(&ctrl->band_lo[cam], lo);
          atomicMax(&ctrl->band_hi[cam], hi);
        }
      }
    }
    __syncthreads();   // table complete; the scratch on the ring is dead from here on

    // bands are copied in 16-byte units: round to multiples of 4 floats (plane % 4 == 0, checked by the host)
    int stage_floats = 4;
    for (int cam = 0; cam < a.n_cam; ++cam) {
      const int lo = ctrl->band_lo[cam] & ~3, hi = (ctrl->band_hi[cam] + 3) & ~3;
      if (hi > lo) stage_floats = max(stage_floats, hi - lo);
    }
    const int n_stages = min(kMaxStages, ring_floats / (CG * stage_floats));   // >= 1: the host sized the ring for CG whole planes
    uint32_t stage = 0, phase = 0;

    if (ctrl->n_entries > tab_cap) {
      // More visible (cell, camera) pairs than the table holds (cameras overlapping almost everywhere): this
      // part is gathered straight from global memory, cell by cell — same arithmetic, no staging.
      if (tid < kConsumers) {
        for (long long item = it; item < seg_end; ++item) {
          const long long r = item % per_part;
          const int b = (int)(r / n_groups);
          const int g0 = (int)(r % n_groups) * CG;
          const int ncg = min(CG, a.C - g0);
          for (int slot = tid; slot < n_cells; slot += kConsumers) {
            const int cell = cell_lo + (int)order[slot];
            float sum[CG];
#pragma unroll
            for (int g = 0; g < CG; ++g) sum[g] = 0.0f;
            int nv = 0;
            for (int cam = 0; cam < a.n_cam; ++cam) {
              float u, v;
              if (!project_cell(a, ctrl->K + cam * 9, ctrl->E + cam * 12, cell % a.W, cell / a.W, u, v)) continue;
              ++nv;
              const Tap tp = make_tap(u, v, a.h, a.w);
              const float* pl = a.feats + (((size_t)b * a.n_cam + cam) * a.C + g0) * plane;
#pragma unroll
              for (int g = 0; g < CG; ++g)
                if (g < ncg) sum[g] = __fadd_rn(sum[g], tap_value(tp, pl + (size_t)g * plane));
            }
            const float den = (float)(nv > 0 ? nv : 1);
#pragma unroll
            for (int g = 0; g < CG; ++g)
              if (g < ncg) a.out[((size_t)b * a.C + g0 + g) * HW + cell] = __fdiv_rn(sum[g], den);
          }
        }
      }
    } else if (warp == kConsumers / 32) {
      // ================================ producer ================================
      // the whole warp runs the loop, one elected lane issues (see elect_one)
      for (long long item = it; item < seg_end; ++item) {
        const long long r = item % per_part;
        const int b = (int)(r / n_groups);
        const int g0 = (int)(r % n_groups) * CG;
        const int ncg = min(CG, a.C - g0);
        for (int cam = 0; cam < a.n_cam; ++cam) {
          const int lo = ctrl->band_lo[cam] & ~3, hi = (ctrl->band_hi[cam] + 3) & ~3;
          if (hi <= lo) continue;   // no cell of the part sees this camera
          const uint32_t bytes = (uint32_t)(hi - lo) * sizeof(float);
          mbarrier_wait(&ctrl->empty[stage], phase ^ 1);
          const float* src = a.feats + (((size_t)b * a.n_cam + cam) * a.C + g0) * plane + lo;
          float* dst = ring + (size_t)stage * CG * stage_floats;
          if (elect_one()) {
            mbarrier_expect_tx(&ctrl->full[stage], bytes * ncg);
#pragma unroll
            for (int g = 0; g < CG; ++g)
              if (g < ncg) bulk_copy_global_to_shared(dst + (size_t)g * stage_floats, src + (size_t)g * plane, bytes, &ctrl->full[stage]);
          }
          __syncwarp();
          if (++stage == (uint32_t)n_stages) { stage = 0; phase ^= 1; }
        }
      }
    } else {
      // ================================ consumers ================================
      float acc[CG][CPT];
#pragma unroll
      for (int g = 0; g < CG; ++g)
#pragma unroll
        for (int j = 0; j < CPT; ++j) acc[g][j] = 0.0f;
      for (long long item = it; item < seg_end; ++item) {
        const long long r = item % per_part;
        const int b = (int)(r / n_groups);
        const int g0 = (int)(r % n_groups) * CG;
        const int ncg = min(CG, a.C - g0);
        for (int cam = 0; cam < a.n_cam; ++cam) {
          const int lo = ctrl->band_lo[cam] & ~3, hi = (ctrl->band_hi[cam] + 3) & ~3;
          if (hi <= lo) continue;
          mbarrier_wait(&ctrl->full[stage], phase);
          const float* sb = ring + (size_t)stage * CG * stage_floats - lo;   // tap offsets are plane-relative
          const uint32_t below = (1u << cam) - 1u;
#pragma unroll
          for (int j = 0; j < CPT; ++j) {
            if ((info[j] >> cam) & 1u) {
              const int idx = (int)(info[j] >> 8) + __popc(info[j] & below);
              const float4 wt = wtab[idx];
              const uint32_t po = otab[idx];
              const float* p00 = sb + (po & 0x3fffffffu);
              const int dx = (int)((po >> 30) & 1u), dy = (po >> 31) ? a.w : 0;
#pragma unroll
              for (int g = 0; g < CG; ++g) {
                const float* pl = p00 + (size_t)g * stage_floats;
                float val = __fmul_rn(wt.x, pl[0]);
                val = __fadd_rn(val, __fmul_rn(wt.y, pl[dx]));
                val = __fadd_rn(val, __fmul_rn(wt.z, pl[dy]));
                val = __fadd_rn(val, __fmul_rn(wt.w, pl[dy + dx]));
                acc[g][j] = __fadd_rn(acc[g][j], val);
              }
            }
          }
          __syncwarp();
          if (lane == 0) mbarrier_arrive(&ctrl->empty[stage]);
          if (++stage == (uint32_t)n_stages) { stage = 0; phase ^= 1; }
        }
        // every canvas value written once; cells of equal visibility are runs of neighbours, so a warp's
        // stores fall into a few contiguous stretches
        float* ob = a.out + ((size_t)b * a.C + g0) * HW;
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
          // mean over the cameras that see the cell: s / n.  n = 1 is the common case and a power of two
          // divides exactly by multiplication; only n = 3, 5, 6, 7 takes the IEEE division.
          float sc = scale[j];
          if (sc == 0.0f) {
            const float den = (float)__popc(info[j] & 0xffu);
#pragma unroll
            for (int g = 0; g < CG; ++g) acc[g][j] = __fdiv_rn(acc[g][j], den);
            sc = 1.0f;
          }
          float* oj = ob + cell_of[j];
#pragma unroll
          for (int g = 0; g < CG; ++g) {
            if (cell_of[j] >= 0 && g < ncg) oj[(size_t)g * HW] = __fmul_rn(acc[g][j], sc);
            acc[g][j] = 0.0f;
          }
        }
      }
    }
    it = seg_end;
    if (it < it_end) __syncthreads();   // the table and the barriers are about to be rebuilt
  }
}

}  // namespace
}  // namespace b200bev

using namespace b200bev;

extern "C" B200BEV_API int b200bev_camera_project(const float* feats, int B, int n_cam, int C, int h, int w, const float* intrinsics,
                                      const float* ego2cam, int T, float img_w, float img_h, float x_min, float y_min,
                                      float voxel_x, float voxel_y, float z_plane, int W, int H, float* out,
                                      float* uv_valid, int impl, void* stream) {
  if (!feats || !intrinsics || !ego2cam || !out || B <= 0 || n_cam <= 0 || C <= 0 || h <= 0 || w <= 0 || W <= 0 || H <= 0)
    return B200BEV_ERR_INVALID_ARGUMENT;
  if (T != 1 && T != B) return B200BEV_ERR_INVALID_ARGUMENT;
  if (impl < B200BEV_PROJECT_AUTO || impl > B200BEV_PROJECT_GATHER) return B200BEV_ERR_INVALID_ARGUMENT;
  if (!(img_w > 0.0f) || !(img_h > 0.0f)) return B200BEV_ERR_INVALID_ARGUMENT;
  if (n_cam > kMaxCams || B > 65535) return B200BEV_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  ProjArgs a{};
  a.feats = feats; a.B = B; a.n_cam = n_cam; a.C = C; a.h = h; a.w = w;
  a.K = intrinsics; a.E = ego2cam; a.T = T;
  a.img_w = img_w; a.img_h = img_h; a.x_min = x_min; a.y_min = y_min; a.vx = voxel_x; a.vy = voxel_y; a.z_plane = z_plane;
  a.W = W; a.H = H; a.out = out; a.uv_valid = uv_valid;
  const long long HW = (long long)H * W;

  if (uv_valid) {
    camera_project_table_kernel<<<dim3((unsigned)ceil_div64(HW, 256), T), 256, 0, st>>>(a);
    B200BEV_CUDA_TRY(cudaGetLastError());
  }

  // staged path: bands are whole 16-byte units.  Shared-memory budget: the table holds up to ~1.25 visible
  // cameras per owned cell (a surround rig overlaps little; pairs beyond that are recomputed on the fly), the
  // ring gets the rest and must hold at least one stage of CG whole planes — bands are normally well under
  // half a plane, which is what gives the ring its depth.
  const long long plane = (long long)h * w;
  const int n_parts = (int)ceil_div64(HW, kPartCells);
  const int part_cells = (int)ceil_div64(HW, n_parts);
  bool staged = (plane % 4 == 0) && ((reinterpret_cast<uintptr_t>(feats) & 15) == 0) && plane < (1LL << 30) && n_parts <= 64;
  int CG = 0, tab_cap = 0, ring_bytes = 0;
  if (staged) {
    const long long want = std::min<long long>((long long)part_cells * n_cam, part_cells + part_cells / 4 + 64);
    tab_cap = (int)((want + 15) / 16 * 16);
    ring_bytes = kMaxSmemOptin - kCtrlBytes - tab_cap * kEntryBytes;
    int cg_max = 4;
    if (const char* e = debug_env("B200BEV_PROJECT_CG")) {   // experiments: 1, 2 or 4
      const int v = atoi(e);
      if (v == 1 || v == 2 || v == 4) cg_max = v;
    }
    const long long scratch = 256 * 4 + kPartCells * 3;   // table-build scratch aliased onto the ring
    for (int cg = cg_max; cg >= 1; cg >>= 1)
      if ((long long)cg * plane * (long long)sizeof(float) <= ring_bytes && scratch <= ring_bytes) { CG = cg; break; }
    if (CG == 0) staged = false;
  }
  if (impl == B200BEV_PROJECT_GATHER) staged = false;
  else if (impl == B200BEV_PROJECT_STAGED && !staged) return B200BEV_ERR_UNSUPPORTED;
  if (staged) {
    const long long items = (long long)n_parts * B * ceil_div(C, CG);
    const int grid = (int)std::min<long long>(items, sm_count());
    const int ring_floats = ring_bytes / (int)sizeof(float);
    auto launch = [&](auto kernel) -> int {
      B200BEV_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemOptin));
      kernel<<<grid, kStagedThreads, kMaxSmemOptin, st>>>(a, n_parts, part_cells, tab_cap, ring_floats);
      return B200BEV_OK;
    };
    const int rc = CG == 4 ? launch(camera_project_staged_kernel<4>) : CG == 2 ? launch(camera_project_staged_kernel<2>)
                                                                              : launch(camera_project_staged_kernel<1>);
    if (rc != B200BEV_OK) return rc;
  } else {
    camera_project_gather_kernel<<<dim3(ceil_div((int)HW, kCellsPerBlock), B), 256, 0, st>>>(a);
  }
  return launch_status();
}
