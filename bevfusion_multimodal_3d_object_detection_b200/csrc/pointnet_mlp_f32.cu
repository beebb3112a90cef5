// S1b/S1c — fused PointNet shared-MLP + max, fp32 (FFMA) path, and the multi-radar fusion head.
//
// Reference arithmetic: PointNetLiDAREncoder.forward (src/encoders.py:289-298) and
// RadarEncoder.forward (src/encoders.py:549-555): x = relu(bn(conv1d_k1(x))) five (four) times, then
// torch.max over the point axis.  The reference materialises every (B, C_l, N) intermediate (143 MB
// per frame for the last one); here a CTA keeps a 64-point tile resident in shared memory through the
// whole layer chain, BatchNorm is folded into the weights on the host (eval mode), and the last
// layer's outputs never leave registers: they are max-reduced over the tile, folded into a running
// per-CTA maximum and flushed with one integer atomicMax per channel per frame (valid because the
// post-ReLU values are >= 0, so their float bit patterns order like integers).
//
//   persistent grid, 256 threads, tile = 64 points.  Per layer and 128-channel block every thread
//   owns an 8(channel) x 4(point) register tile; activations sit in smem as [channel][point], the
//   W^T tiles ([k][128 channels], 16 k at a time) are double-buffered with cp.async.
//
// Cell mode (perm/offsets from b200bev_bin_sort): tiles walk the points in cell order, the last
// layer is max-reduced per run of equal cells and merged into the channels-last canvas with
// atomicMax — the per-cell scatter-max north_star describes.
#include "common.cuh"

namespace b200bev {
namespace {

constexpr int kP = 64;        // points per tile
constexpr int kNB = 128;      // output channels per block pass
constexpr int kKC = 16;       // k rows per weight stage
constexpr int kThreads = 256;
constexpr int kMaxLayers = 8;
constexpr int kMaxRadars = 8;

struct MlpArgs {
  const float* pts;  // lidar mode: (B,N,C)
  int B, N, C;
  const float* params;
  int n_layers;
  int dims[kMaxLayers + 1];
  long long w_off[kMaxLayers], b_off[kMaxLayers];
  const int32_t* perm;
  const int32_t* offsets;
  int n_cells;
  float* out;     // global maxima (frames, C_out) or nullptr
  float* canvas;  // per-cell maxima (frames, n_cells, C_out) or nullptr
  int frames, tiles_per_frame, rows_a, rows_b;
  // radar mode (R > 0): frame f = b*R + r reads radar r
  int R;
  const float* radar_pts[kMaxRadars];
  int radar_n[kMaxRadars];
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, bool pred) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  const int sz = pred ? 16 : 0;  // src-size 0 -> 16 bytes of zeros, nothing is read
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gmem_src), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N));
}

// stage rows [k0, k0+kKC) x channels [n0, n0+kNB) of W^T (K x Nout, row-major) into wbuf[kk*kNB + n]
__device__ __forceinline__ void stage_weights(float* wbuf, const float* __restrict__ wt, int K, int Nout, int k0, int n0) {
  for (int ch = threadIdx.x; ch < kKC * (kNB / 4); ch += kThreads) {
    const int kk = ch / (kNB / 4), n = (ch % (kNB / 4)) * 4;
    const bool ok = (k0 + kk < K) && (n0 + n < Nout);
    cp_async16(wbuf + kk * kNB + n, ok ? (wt + (size_t)(k0 + kk) * Nout + n0 + n) : wt, ok);
  }
}

__global__ void __launch_bounds__(kThreads, 1) pointnet_mlp_f32_kernel(MlpArgs a) {
  extern __shared__ __align__(16) float smem_f[];
  float* buf_a = smem_f;                          // rows_a x kP
  float* buf_b = buf_a + (size_t)a.rows_a * kP;   // rows_b x kP
  float* wbuf = buf_b + (size_t)a.rows_b * kP;    // 2 x kKC x kNB
  float* runmax = wbuf + 2 * kKC * kNB;           // dims[n_layers]
  int* cellid = reinterpret_cast<int*>(runmax + a.dims[a.n_layers]);  // kP
  int* pidx = cellid + kP;                                           // kP: source row of each tile slot, -1 = none

  const int tid = threadIdx.x;
  const int tp = tid & 15, tn = tid >> 4;
  const int c_out = a.dims[a.n_layers];
  const bool sorted = a.perm != nullptr;   // walk the points in cell order
  const bool want_global = a.out != nullptr;
  const bool want_canvas = a.canvas != nullptr;

  const long long total = (long long)a.frames * a.tiles_per_frame;
  // contiguous shares that differ by at most one tile (the first `extra` CTAs take one more)
  const long long base_share = total / gridDim.x, extra = total % gridDim.x;
  const long long t_begin = base_share * blockIdx.x + (blockIdx.x < extra ? blockIdx.x : extra);
  const long long t_end = t_begin + base_share + (blockIdx.x < extra ? 1 : 0);

  for (int n = tid; n < c_out; n += kThreads) runmax[n] = 0.0f;
  int cur_frame = -1;

  for (long long t = t_begin; t < t_end; ++t) {
    const int f = (int)(t / a.tiles_per_frame);
    const int tile = (int)(t % a.tiles_per_frame);
    // frame geometry
    const float* fpts;
    int n_pts;
    if (a.R > 0) {
      const int r = f % a.R, b = f / a.R;
      n_pts = a.radar_n[r];
      fpts = a.radar_pts[r] + (size_t)b * n_pts * a.C;
    } else {
      n_pts = a.N;
      fpts = a.pts + (size_t)f * a.N * a.C;
    }
    int n_walk = n_pts, n_in_grid = 0;
    const int32_t* fperm = nullptr;
    const int32_t* foff = nullptr;
    if (sorted) {
      fperm = a.perm + (size_t)f * a.N;
      foff = a.offsets + (size_t)f * (a.n_cells + 1);
      n_in_grid = __ldg(foff + a.n_cells);
      if (!want_global) n_walk = n_in_grid;  // the out-of-grid tail only matters for the global max
    }
    const int s0 = tile * kP;
    if (s0 >= n_walk) continue;  // uniform across the CTA

    if (want_global && f != cur_frame) {
      __syncthreads();
      if (cur_frame >= 0) {
        int* o = reinterpret_cast<int*>(a.out + (size_t)cur_frame * c_out);
        for (int n = tid; n < c_out; n += kThreads) {
          atomicMax(o + n, __float_as_int(runmax[n]));
          runmax[n] = 0.0f;
        }
      }
      cur_frame = f;
    }
    __syncthreads();  // previous tile fully consumed before its buffers are overwritten

    // ---- tile slots -> source rows (+ cell ids) ----
    if (tid < kP) {
      const int s = s0 + tid;
      int src = -1, cid = -1;
      if (s < n_walk) {
        if (sorted) {
          src = __ldg(fperm + s);
          if (s < n_in_grid) {
            int lo = 0, hi = a.n_cells;  // largest c with offsets[c] <= s
            while (hi - lo > 1) {
              const int mid = (lo + hi) >> 1;
              if (__ldg(foff + mid) <= s) lo = mid; else hi = mid;
            }
            cid = lo;
          }
        } else {
          src = s;
        }
      }
      pidx[tid] = src;
      cellid[tid] = cid;
    }
    __syncthreads();
    // ---- layer-0 input: act[k][p] = pts[row(p)][k] ----
    for (int e = tid; e < a.C * kP; e += kThreads) {
      const int p = e / a.C, k = e % a.C;
      const int src = pidx[p];
      buf_a[k * kP + p] = src >= 0 ? __ldg(fpts + (size_t)src * a.C + k) : 0.0f;
    }
    __syncthreads();

    for (int l = 0; l < a.n_layers; ++l) {
      const int K = a.dims[l], Nout = a.dims[l + 1];
      const float* __restrict__ wt = a.params + a.w_off[l];
      const float* __restrict__ bias = a.params + a.b_off[l];
      const float* in = (l & 1) ? buf_b : buf_a;
      float* outb = (l & 1) ? buf_a : buf_b;
      const bool last = (l == a.n_layers - 1);
      const int nk = ceil_div(K, kKC);

      for (int n0 = 0; n0 < Nout; n0 += kNB) {
        float acc[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[j][i] = 0.0f;

        stage_weights(wbuf, wt, K, Nout, 0, n0);
        cp_async_commit();
        for (int kc = 0; kc < nk; ++kc) {
          if (kc + 1 < nk) {
            stage_weights(wbuf + ((kc + 1) & 1) * kKC * kNB, wt, K, Nout, (kc + 1) * kKC, n0);
            cp_async_commit();
            cp_async_wait<1>();
          } else {
            cp_async_wait<0>();
          }
          __syncthreads();
          const float* w = wbuf + (kc & 1) * kKC * kNB + tn * 8;
          const float* x = in + (size_t)kc * kKC * kP + tp * 4;
          const int kmax = min(kKC, K - kc * kKC);
          if (kmax == kKC) {
#pragma unroll
            for (int kk = 0; kk < kKC; ++kk) {
              const float4 xv = *reinterpret_cast<const float4*>(x + kk * kP);
              const float4 w0 = *reinterpret_cast<const float4*>(w + kk * kNB);
              const float4 w1 = *reinterpret_cast<const float4*>(w + kk * kNB + 4);
              const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
              const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
              for (int j = 0; j < 8; ++j)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[j][i] = fmaf(wv[j], xs[i], acc[j][i]);
            }
          } else {
            for (int kk = 0; kk < kmax; ++kk) {
              const float4 xv = *reinterpret_cast<const float4*>(x + kk * kP);
              const float4 w0 = *reinterpret_cast<const float4*>(w + kk * kNB);
              const float4 w1 = *reinterpret_cast<const float4*>(w + kk * kNB + 4);
              const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
              const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
              for (int j = 0; j < 8; ++j)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[j][i] = fmaf(wv[j], xs[i], acc[j][i]);
            }
          }
          __syncthreads();
        }

        const int nbase = n0 + tn * 8;
        if (nbase < Nout) {  // widths are multiples of 8, so the 8-channel tile is all-in or all-out
          if (!last) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float bj = __ldg(bias + nbase + j);
              float4 v;
              v.x = fmaxf(acc[j][0] + bj, 0.0f);
              v.y = fmaxf(acc[j][1] + bj, 0.0f);
              v.z = fmaxf(acc[j][2] + bj, 0.0f);
              v.w = fmaxf(acc[j][3] + bj, 0.0f);
              *reinterpret_cast<float4*>(outb + (size_t)(nbase + j) * kP + tp * 4) = v;
            }
          } else {
            bool ok[4];
            int cid[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              ok[i] = pidx[tp * 4 + i] >= 0;
              cid[i] = cellid[tp * 4 + i];
            }
            float* canvas = want_canvas ? a.canvas + (size_t)f * a.n_cells * c_out : nullptr;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float bj = __ldg(bias + nbase + j);
              float v[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) v[i] = acc[j][i] + bj;
              if (want_canvas) {
                float m = 0.0f;  // relu floor
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  if (cid[i] >= 0) {
                    m = fmaxf(m, v[i]);
                    const bool run_end = (i == 3) || (cid[i + 1 < 4 ? i + 1 : 3] != cid[i]);
                    if (run_end) {
                      if (m > 0.0f)
                        atomicMax(reinterpret_cast<int*>(canvas + (size_t)cid[i] * c_out + nbase + j), __float_as_int(m));
                      m = 0.0f;
                    }
                  }
                }
              }
              if (want_global) {
                float m = 0.0f;  // relu floor; empty slots contribute nothing
#pragma unroll
                for (int i = 0; i < 4; ++i)
                  if (ok[i]) m = fmaxf(m, v[i]);
                // the two half-warps hold different channel tiles and may diverge on `nbase < Nout`
                // (widths that are multiples of 8 but not 16): reduce within the half-warp only
                const unsigned half = (tid & 16) ? 0xffff0000u : 0x0000ffffu;
                m = fmaxf(m, __shfl_xor_sync(half, m, 1));
                m = fmaxf(m, __shfl_xor_sync(half, m, 2));
                m = fmaxf(m, __shfl_xor_sync(half, m, 4));
                m = fmaxf(m, __shfl_xor_sync(half, m, 8));
                if (tp == 0) runmax[nbase + j] = fmaxf(runmax[nbase + j], m);
              }
            }
          }
        }
      }
      __syncthreads();
    }
  }
  if (want_global) {
    __syncthreads();
    if (cur_frame >= 0) {
      int* o = reinterpret_cast<int*>(a.out + (size_t)cur_frame * c_out);
      for (int n = tid; n < c_out; n += kThreads) atomicMax(o + n, __float_as_int(runmax[n]));
    }
  }
}

// Multi-radar fusion (src/encoders.py:647-659): concat -> Linear, or max / mean over radars.
// One warp per output element (b, o).
__global__ void __launch_bounds__(256) radar_fuse_kernel(const float* __restrict__ per_radar, int B, int R, int F,
                                                         int fusion, const float* __restrict__ w,
                                                         const float* __restrict__ bias, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (gw >= B * F) return;
  const int b = gw / F, o = gw % F;
  const float* x = per_radar + (size_t)b * R * F;
  if (fusion == B200BEV_RADAR_CONCAT) {
    const int n = R * F;
    const float* wr = w + (size_t)o * n;
    float s = 0.0f;
    for (int i = lane; i < n; i += 32) s = fmaf(__ldg(wr + i), __ldg(x + i), s);
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(FULL_MASK, s, d);
    if (lane == 0) out[gw] = s + __ldg(bias + o);
  } else if (lane == 0) {
    float s = __ldg(x + o);
    for (int r = 1; r < R; ++r) {
      const float v = __ldg(x + (size_t)r * F + o);
      s = (fusion == B200BEV_RADAR_MAX) ? fmaxf(s, v) : s + v;
    }
    out[gw] = (fusion == B200BEV_RADAR_MAX) ? s : __fdiv_rn(s, (float)R);
  }
}

int fill_layers(MlpArgs& a, const int32_t* dims, int n_layers, int C) {
  if (!dims || n_layers < 1 || n_layers > kMaxLayers || dims[0] != C) return B200BEV_ERR_INVALID_ARGUMENT;
  a.n_layers = n_layers;
  long long off = 0;
  int rows_a = 0, rows_b = 0;
  for (int l = 0; l <= n_layers; ++l) {
    a.dims[l] = dims[l];
    if (dims[l] <= 0 || dims[l] > 2048) return B200BEV_ERR_UNSUPPORTED;
    if (l > 0 && (dims[l] % 8) != 0) return B200BEV_ERR_UNSUPPORTED;
    if (l < n_layers) {  // the last layer's output is never stored
      if (l & 1) rows_b = dims[l] > rows_b ? dims[l] : rows_b;
      else rows_a = dims[l] > rows_a ? dims[l] : rows_a;
    }
  }
  for (int l = 0; l < n_layers; ++l) {
    a.w_off[l] = off;
    off += (long long)dims[l] * dims[l + 1];
    a.b_off[l] = off;
    off += dims[l + 1];
  }
  a.rows_a = rows_a;
  a.rows_b = rows_b > 0 ? rows_b : 1;
  return B200BEV_OK;
}

size_t mlp_smem_bytes(const MlpArgs& a) {
  return ((size_t)(a.rows_a + a.rows_b) * kP + 2 * kKC * kNB + a.dims[a.n_layers]) * sizeof(float) + 2 * kP * sizeof(int);
}

int launch_mlp(MlpArgs& a, cudaStream_t st) {
  const size_t smem = mlp_smem_bytes(a);
  if (smem > 227 * 1024) return B200BEV_ERR_UNSUPPORTED;
  if (smem > 48 * 1024)
    B200BEV_CUDA_TRY(cudaFuncSetAttribute(pointnet_mlp_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long total = (long long)a.frames * a.tiles_per_frame;
  // persistent grid: as many CTAs as can be resident at once (the narrow radar MLP fits three per SM; with one per
  // SM its 320 tiles took three rounds on 107 SMs), each walking a contiguous share of the tiles
  int resident = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, pointnet_mlp_f32_kernel, kThreads, smem) != cudaSuccess || resident < 1)
    resident = 1;
  long long grid = (long long)sm_count() * resident;
  if (grid > total) grid = total;
  pointnet_mlp_f32_kernel<<<(int)grid, kThreads, smem, st>>>(a);
  return launch_status();
}

}  // namespace

// fp32 path of b200bev_pointnet_encode (dispatch lives in api.cu)
int pointnet_encode_f32(const float* points, int B, int N, int C, const float* params, const int32_t* dims,
                        int n_layers, const int32_t* perm, const int32_t* offsets, int n_cells, float* out_global,
                        float* out_canvas, cudaStream_t st) {
  MlpArgs a{};
  const int rc = fill_layers(a, dims, n_layers, C);
  if (rc) return rc;
  if ((reinterpret_cast<uintptr_t>(params) & 15) != 0) return B200BEV_ERR_INVALID_ARGUMENT;
  a.pts = points; a.B = B; a.N = N; a.C = C; a.params = params;
  a.perm = perm; a.offsets = offsets; a.n_cells = n_cells; a.out = out_global; a.canvas = out_canvas;
  a.frames = B;
  a.tiles_per_frame = ceil_div(N, kP);
  a.R = 0;
  const size_t c_out = (size_t)dims[n_layers];
  if (out_global) B200BEV_CUDA_TRY(cudaMemsetAsync(out_global, 0, (size_t)B * c_out * sizeof(float), st));
  if (out_canvas) B200BEV_CUDA_TRY(cudaMemsetAsync(out_canvas, 0, (size_t)B * n_cells * c_out * sizeof(float), st));
  return launch_mlp(a, st);
}

}  // namespace b200bev

using namespace b200bev;

extern "C" B200BEV_API int b200bev_radar_encode(const float* const* radar_points, const int32_t* n_points, int R, int B, int C,
                                    const float* params, const int32_t* dims, int n_layers, int fusion,
                                    const float* fc_weight, const float* fc_bias, float* per_radar, float* out,
                                    void* stream) {
  if (!radar_points || !n_points || !params || !per_radar || !out || R <= 0 || B <= 0 || C <= 0)
    return B200BEV_ERR_INVALID_ARGUMENT;
  if (R > kMaxRadars) return B200BEV_ERR_UNSUPPORTED;
  if (fusion < B200BEV_RADAR_CONCAT || fusion > B200BEV_RADAR_MEAN) return B200BEV_ERR_INVALID_ARGUMENT;
  if (fusion == B200BEV_RADAR_CONCAT && (!fc_weight || !fc_bias)) return B200BEV_ERR_INVALID_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  MlpArgs a{};
  const int rc = fill_layers(a, dims, n_layers, C);
  if (rc) return rc;
  if ((reinterpret_cast<uintptr_t>(params) & 15) != 0) return B200BEV_ERR_INVALID_ARGUMENT;
  int max_n = 0;
  for (int r = 0; r < R; ++r) {
    if (!radar_points[r] || n_points[r] <= 0) return B200BEV_ERR_INVALID_ARGUMENT;  // torch.max over 0 points raises
    a.radar_pts[r] = radar_points[r];
    a.radar_n[r] = n_points[r];
    max_n = n_points[r] > max_n ? n_points[r] : max_n;
  }
  a.R = R; a.B = B; a.N = max_n; a.C = C; a.params = params; a.out = per_radar; a.canvas = nullptr;
  a.frames = B * R;
  a.tiles_per_frame = ceil_div(max_n, kP);
  const int F = dims[n_layers];
  B200BEV_CUDA_TRY(cudaMemsetAsync(per_radar, 0, (size_t)B * R * F * sizeof(float), st));
  const int rc2 = launch_mlp(a, st);
  if (rc2) return rc2;
  const long long warps = (long long)B * F;
  radar_fuse_kernel<<<(int)((warps * 32 + 255) / 256), 256, 0, st>>>(per_radar, B, R, F, fusion, fc_weight, fc_bias, out);
  return launch_status();
}
