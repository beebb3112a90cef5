// S2 — camera features -> BEV canvas.
//
//   camera_mean      camera_features.mean(dim=1)                       src/fusion.py:233-234
//   bilinear_resize  F.interpolate(size, 'bilinear', align_corners=False)  src/fusion.py:242-247
//   camera_project   BEV cell centres -> pinhole projection -> bilinear gather -> mean over the
//                    cameras that see the cell (north_star S2; the reference has no geometric
//                    projection, SURVEY §0 — oracle/bev_oracle.py holds the restatement)
//
// All three are HBM-bound streaming/gather kernels: 16-byte accesses where the layout allows,
// every output written exactly once, no intermediate tensors.
#include "common.cuh"

namespace b200bev {
namespace {

// ---------------------------------------------------------------------------------------------
// mean over cameras: out[b, j] = (f[b,0,j] + f[b,1,j] + ... + f[b,n-1,j]) / n, summed in camera
// order (the association torch's CPU sum uses for a 6-long reduction), IEEE divide.
// ---------------------------------------------------------------------------------------------
template <int NCAM>
__global__ void __launch_bounds__(256) camera_mean_vec4_kernel(const float4* __restrict__ f, float4* __restrict__ out,
                                                               int n_cam_rt, long long inner4, long long total4) {
  const int n_cam = NCAM > 0 ? NCAM : n_cam_rt;
  const float denom = (float)n_cam;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4;
       i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / inner4, j = i - b * inner4;
    const float4* src = f + b * n_cam * inner4 + j;
    float4 v[NCAM > 0 ? NCAM : 1];
    float4 s;
    if (NCAM > 0) {
#pragma unroll
      for (int c = 0; c < NCAM; ++c) v[c] = ld_stream_f4(src + (long long)c * inner4);  // all loads in flight
      s = v[0];
#pragma unroll
      for (int c = 1; c < NCAM; ++c) {
        s.x = __fadd_rn(s.x, v[c].x); s.y = __fadd_rn(s.y, v[c].y);
        s.z = __fadd_rn(s.z, v[c].z); s.w = __fadd_rn(s.w, v[c].w);
      }
    } else {
      s = ld_stream_f4(src);
      for (int c = 1; c < n_cam; ++c) {
        const float4 t = ld_stream_f4(src + (long long)c * inner4);
        s.x = __fadd_rn(s.x, t.x); s.y = __fadd_rn(s.y, t.y);
        s.z = __fadd_rn(s.z, t.z); s.w = __fadd_rn(s.w, t.w);
      }
    }
    s.x = __fdiv_rn(s.x, denom); s.y = __fdiv_rn(s.y, denom);
    s.z = __fdiv_rn(s.z, denom); s.w = __fdiv_rn(s.w, denom);
    out[i] = s;
  }
}

__global__ void __launch_bounds__(256) camera_mean_scalar_kernel(const float* __restrict__ f, float* __restrict__ out,
                                                                 int n_cam, long long inner, long long total) {
  const float denom = (float)n_cam;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / inner, j = i - b * inner;
    const float* src = f + b * n_cam * inner + j;
    float s = __ldg(src);
    for (int c = 1; c < n_cam; ++c) s = __fadd_rn(s, __ldg(src + (long long)c * inner));
    out[i] = __fdiv_rn(s, denom);
  }
}

// ---------------------------------------------------------------------------------------------
// bilinear resize, align_corners=False (aten upsample_bilinear2d semantics):
//   scale = in/out;  src = max(scale*(dst+0.5) - 0.5, 0);  i0 = min(floor(src), in-1);
//   i1 = min(i0+1, in-1);  l1 = clamp(src - i0, 0, 1);  l0 = 1 - l1
//   out = l0y*(l0x*v00 + l1x*v01) + l1y*(l0x*v10 + l1x*v11)
// One thread per output pixel of one plane; x fastest so writes coalesce and the four taps of
// neighbouring threads fall in the same input rows.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void resize_coord(int dst, float scale, int in_size, int& i0, int& i1, float& l0, float& l1) {
  float src = __fsub_rn(__fmul_rn(scale, __fadd_rn((float)dst, 0.5f)), 0.5f);
  if (src < 0.0f) src = 0.0f;
  i0 = min((int)floorf(src), in_size - 1);
  i1 = min(i0 + 1, in_size - 1);
  l1 = fminf(fmaxf(__fsub_rn(src, (float)i0), 0.0f), 1.0f);
  l0 = __fsub_rn(1.0f, l1);
}

__global__ void __launch_bounds__(256) bilinear_resize_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                              long long planes, int h, int w, int H, int W) {
  const float sy = __fdiv_rn((float)h, (float)H), sx = __fdiv_rn((float)w, (float)W);
  const long long total = planes * H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    const long long t = i / W;
    const int y = (int)(t % H);
    const long long p = t / H;
    int y0, y1, x0, x1;
    float ly0, ly1, lx0, lx1;
    resize_coord(y, sy, h, y0, y1, ly0, ly1);
    resize_coord(x, sx, w, x0, x1, lx0, lx1);
    const float* src = in + p * h * w;
    const float v00 = __ldg(src + y0 * w + x0), v01 = __ldg(src + y0 * w + x1);
    const float v10 = __ldg(src + y1 * w + x0), v11 = __ldg(src + y1 * w + x1);
    const float top = __fadd_rn(__fmul_rn(lx0, v00), __fmul_rn(lx1, v01));
    const float bot = __fadd_rn(__fmul_rn(lx0, v10), __fmul_rn(lx1, v11));
    out[i] = __fadd_rn(__fmul_rn(ly0, top), __fmul_rn(ly1, bot));
  }
}

// ---------------------------------------------------------------------------------------------
// geometric projection + gather.
// ---------------------------------------------------------------------------------------------
constexpr int kMaxCams = 8;
constexpr int kCellsPerBlock = 64;
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

// Feature-map coordinates (u,v) of BEV cell (ix,iy) in camera `cam`; every operation is a single
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
  int o00, o01, o10, o11;   // clamped offsets inside one (h,w) plane
  float w00, w01, w10, w11; // bilinear weights, 0 for taps outside the map (padding_mode='zeros')
};

__global__ void __launch_bounds__(256) camera_project_kernel(ProjArgs a) {
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
        const float* Kc = a.K + ((size_t)t * a.n_cam + cam) * 9;
        const float* Ec = a.E + ((size_t)t * a.n_cam + cam) * 12;
        float u, v;
        const bool vis = project_cell(a, Kc, Ec, ix, iy, u, v);
        if (a.uv_valid && (a.T != 1 || b == 0)) {
          float* o = a.uv_valid + (((size_t)t * HW + cell) * a.n_cam + cam) * 3;
          o[0] = u; o[1] = v; o[2] = vis ? 1.0f : 0.0f;
        }
        if (vis) {
          const float fx = floorf(u), fy = floorf(v);
          const int x0 = (int)fx, y0 = (int)fy;
          const float ax = __fsub_rn(u, fx), ay = __fsub_rn(v, fy);  // weight of the +1 tap
          const float bx = __fsub_rn(__fadd_rn(fx, 1.0f), u), by = __fsub_rn(__fadd_rn(fy, 1.0f), v);
          const bool x0ok = x0 >= 0 && x0 < a.w, x1ok = x0 + 1 >= 0 && x0 + 1 < a.w;
          const bool y0ok = y0 >= 0 && y0 < a.h, y1ok = y0 + 1 >= 0 && y0 + 1 < a.h;
          const int cx0 = min(max(x0, 0), a.w - 1), cx1 = min(max(x0 + 1, 0), a.w - 1);
          const int cy0 = min(max(y0, 0), a.h - 1), cy1 = min(max(y0 + 1, 0), a.h - 1);
          Tap tp;
          tp.o00 = cy0 * a.w + cx0; tp.o01 = cy0 * a.w + cx1;
          tp.o10 = cy1 * a.w + cx0; tp.o11 = cy1 * a.w + cx1;
          tp.w00 = (x0ok && y0ok) ? __fmul_rn(bx, by) : 0.0f;
          tp.w01 = (x1ok && y0ok) ? __fmul_rn(ax, by) : 0.0f;
          tp.w10 = (x0ok && y1ok) ? __fmul_rn(bx, ay) : 0.0f;
          tp.w11 = (x1ok && y1ok) ? __fmul_rn(ax, ay) : 0.0f;
          taps[tid][nv] = tp;
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
  const float inv_den = (float)(nv > 0 ? nv : 1);
  for (int c = cgroup; c < a.C; c += kGroups) {
    float s = 0.0f;
    for (int k = 0; k < nv; ++k) {
      const Tap& tp = taps[lane_cell][k];
      const float* src = fb + ((size_t)cam_of[lane_cell][k] * a.C + c) * plane;
      float val = __fmul_rn(tp.w00, __ldg(src + tp.o00));
      val = __fadd_rn(val, __fmul_rn(tp.w01, __ldg(src + tp.o01)));
      val = __fadd_rn(val, __fmul_rn(tp.w10, __ldg(src + tp.o10)));
      val = __fadd_rn(val, __fmul_rn(tp.w11, __ldg(src + tp.o11)));
      s = __fadd_rn(s, val);
    }
    ob[(size_t)c * HW] = __fdiv_rn(s, inv_den);
  }
}

int grid_for(long long work_items, int per_block) {
  long long blocks = (work_items + per_block - 1) / per_block;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace
}  // namespace b200bev

using namespace b200bev;

extern "C" B200BEV_API int b200bev_camera_mean(const float* feats, int B, int n_cam, int64_t inner, float* out, void* stream) {
  if (!feats || !out || B <= 0 || n_cam <= 0 || inner <= 0) return B200BEV_ERR_INVALID_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = (inner % 4 == 0) && ((reinterpret_cast<uintptr_t>(feats) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  if (vec) {
    const long long inner4 = inner / 4, total4 = (long long)B * inner4;
    const int grid = grid_for(total4, 256);
    const float4* f4 = reinterpret_cast<const float4*>(feats);
    float4* o4 = reinterpret_cast<float4*>(out);
    if (n_cam == 6) camera_mean_vec4_kernel<6><<<grid, 256, 0, st>>>(f4, o4, n_cam, inner4, total4);
    else camera_mean_vec4_kernel<0><<<grid, 256, 0, st>>>(f4, o4, n_cam, inner4, total4);
  } else {
    const long long total = (long long)B * inner;
    camera_mean_scalar_kernel<<<grid_for(total, 256), 256, 0, st>>>(feats, out, n_cam, inner, total);
  }
  return launch_status();
}

extern "C" B200BEV_API int b200bev_bilinear_resize(const float* in, int B, int C, int h, int w, float* out, int H, int W, void* stream) {
  if (!in || !out || B <= 0 || C <= 0 || h <= 0 || w <= 0 || H <= 0 || W <= 0) return B200BEV_ERR_INVALID_ARGUMENT;
  const long long planes = (long long)B * C;
  bilinear_resize_kernel<<<grid_for(planes * H * W, 256), 256, 0, (cudaStream_t)stream>>>(in, out, planes, h, w, H, W);
  return launch_status();
}

extern "C" B200BEV_API int b200bev_camera_project(const float* feats, int B, int n_cam, int C, int h, int w, const float* intrinsics,
                                      const float* ego2cam, int T, float img_w, float img_h, float x_min, float y_min,
                                      float voxel_x, float voxel_y, float z_plane, int W, int H, float* out,
                                      float* uv_valid, void* stream) {
  if (!feats || !intrinsics || !ego2cam || !out || B <= 0 || n_cam <= 0 || C <= 0 || h <= 0 || w <= 0 || W <= 0 || H <= 0)
    return B200BEV_ERR_INVALID_ARGUMENT;
  if (T != 1 && T != B) return B200BEV_ERR_INVALID_ARGUMENT;
  if (!(img_w > 0.0f) || !(img_h > 0.0f)) return B200BEV_ERR_INVALID_ARGUMENT;
  if (n_cam > kMaxCams || B > 65535) return B200BEV_ERR_UNSUPPORTED;
  ProjArgs a{};
  a.feats = feats; a.B = B; a.n_cam = n_cam; a.C = C; a.h = h; a.w = w;
  a.K = intrinsics; a.E = ego2cam; a.T = T;
  a.img_w = img_w; a.img_h = img_h; a.x_min = x_min; a.y_min = y_min; a.vx = voxel_x; a.vy = voxel_y; a.z_plane = z_plane;
  a.W = W; a.H = H; a.out = out; a.uv_valid = uv_valid;
  camera_project_kernel<<<dim3(ceil_div(H * W, kCellsPerBlock), B), 256, 0, (cudaStream_t)stream>>>(a);
  return launch_status();
}
