// S2 — camera features -> BEV canvas.
//
//   camera_mean      camera_features.mean(dim=1)                       src/fusion.py:233-234
//   bilinear_resize  F.interpolate(size, 'bilinear', align_corners=False)  src/fusion.py:242-247
//   (camera_project, the geometric form, lives in camera_project.cu)
//
// All three are HBM-bound streaming/gather kernels: 16-byte accesses where the layout allows,
// every output written exactly once, no intermediate tensors.
#include <cuda_bf16.h>

#include <algorithm>

#include "async_copy.cuh"
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
  const bool small = total < (1ll << 31);     // 32-bit index arithmetic: two 64-bit divisions per output were most of this kernel
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int x, y;
    long long p;
    if (small) {
      const unsigned iu = (unsigned)i, tu = iu / (unsigned)W, pu = tu / (unsigned)H;
      x = (int)(iu - tu * (unsigned)W);
      y = (int)(tu - pu * (unsigned)H);
      p = pu;
    } else {
      x = (int)(i % W);
      const long long t = i / W;
      y = (int)(t % H);
      p = t / H;
    }
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

// The same arithmetic on channels-last bf16 (the tensors between the tcgen05 convolution kernels): in (B,h,w,C) bf16 ->
// channels [c_offset, c_offset + C) of (B,H,W,C_total) bf16.  A thread owns one output pixel and eight channels: four 16-byte
// loads (the taps), fp32 interpolation in the op order above, one 16-byte store; the channel groups of a pixel are
// neighbouring threads, so every access of a warp is contiguous.
__global__ void __launch_bounds__(256) bilinear_resize_nhwc_bf16_kernel(const __nv_bfloat16* __restrict__ in, int B, int h, int w, int C,
                                                                        __nv_bfloat16* __restrict__ out, int H, int W, int C_total, int c_offset) {
  const float sy = __fdiv_rn((float)h, (float)H), sx = __fdiv_rn((float)w, (float)W);
  const unsigned cg = (unsigned)C / 8;
  const long long total = (long long)B * H * W * cg;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const unsigned g = (unsigned)(i % cg);
    const long long pix = i / cg;
    const int x = (int)(pix % W);
    const long long t = pix / W;
    const int y = (int)(t % H), b = (int)(t / H);
    int y0, y1, x0, x1;
    float ly0, ly1, lx0, lx1;
    resize_coord(y, sy, h, y0, y1, ly0, ly1);
    resize_coord(x, sx, w, x0, x1, lx0, lx1);
    const __nv_bfloat16* src = in + (size_t)b * h * w * C + g * 8;
    const uint4 q00 = __ldg(reinterpret_cast<const uint4*>(src + ((size_t)y0 * w + x0) * C));
    const uint4 q01 = __ldg(reinterpret_cast<const uint4*>(src + ((size_t)y0 * w + x1) * C));
    const uint4 q10 = __ldg(reinterpret_cast<const uint4*>(src + ((size_t)y1 * w + x0) * C));
    const uint4 q11 = __ldg(reinterpret_cast<const uint4*>(src + ((size_t)y1 * w + x1) * C));
    const __nv_bfloat162* a00 = reinterpret_cast<const __nv_bfloat162*>(&q00);
    const __nv_bfloat162* a01 = reinterpret_cast<const __nv_bfloat162*>(&q01);
    const __nv_bfloat162* a10 = reinterpret_cast<const __nv_bfloat162*>(&q10);
    const __nv_bfloat162* a11 = reinterpret_cast<const __nv_bfloat162*>(&q11);
    __align__(16) __nv_bfloat162 r[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 v00 = __bfloat1622float2(a00[e]), v01 = __bfloat1622float2(a01[e]);
      const float2 v10 = __bfloat1622float2(a10[e]), v11 = __bfloat1622float2(a11[e]);
      const float tx = __fadd_rn(__fmul_rn(lx0, v00.x), __fmul_rn(lx1, v01.x)), bx = __fadd_rn(__fmul_rn(lx0, v10.x), __fmul_rn(lx1, v11.x));
      const float ty = __fadd_rn(__fmul_rn(lx0, v00.y), __fmul_rn(lx1, v01.y)), by = __fadd_rn(__fmul_rn(lx0, v10.y), __fmul_rn(lx1, v11.y));
      r[e] = __floats2bfloat162_rn(__fadd_rn(__fmul_rn(ly0, tx), __fmul_rn(ly1, bx)), __fadd_rn(__fmul_rn(ly0, ty), __fmul_rn(ly1, by)));
    }
    *reinterpret_cast<uint4*>(out + (size_t)pix * C_total + c_offset + g * 8) = *reinterpret_cast<const uint4*>(r);
  }
}

// Staged form of the same arithmetic: a persistent CTA double-buffers whole input planes in shared
// memory with cp.async.bulk (one elected thread issues, an mbarrier counts the bytes), so HBM sees
// only sequential full-line reads; the taps come out of shared memory, the x/y coordinate tables are
// computed once per CTA, and every output row is written coalesced.
struct ResizeAxis {
  int i0, i1;
  float l0, l1;
};

__global__ void __launch_bounds__(256) bilinear_resize_staged_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                                     long long planes, int h, int w, int H, int W) {
  extern __shared__ __align__(128) uint8_t rs_smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(rs_smem);                 // [2]
  ResizeAxis* xtab = reinterpret_cast<ResizeAxis*>(rs_smem + 16);        // [W]
  ResizeAxis* ytab = xtab + W;                                           // [H]
  const int plane = h * w;
  float* buf = reinterpret_cast<float*>(rs_smem + ((16 + (size_t)(W + H) * sizeof(ResizeAxis) + 127) & ~(size_t)127));  // [2][plane]
  const int tid = threadIdx.x;
  const float sy = __fdiv_rn((float)h, (float)H), sx = __fdiv_rn((float)w, (float)W);
  for (int i = tid; i < W; i += blockDim.x) resize_coord(i, sx, w, xtab[i].i0, xtab[i].i1, xtab[i].l0, xtab[i].l1);
  for (int i = tid; i < H; i += blockDim.x) {
    ResizeAxis r;
    resize_coord(i, sy, h, r.i0, r.i1, r.l0, r.l1);
    r.i0 *= w;   // row offsets
    r.i1 *= w;
    ytab[i] = r;
  }
  if (tid == 0) {
    mbarrier_init(&full[0], 1);
    mbarrier_init(&full[1], 1);
    mbarrier_init_fence();
  }
  __syncthreads();
  const uint32_t bytes = (uint32_t)plane * sizeof(float);
  long long p = blockIdx.x;
  // copies are issued by an elected lane of the (converged) first warp: see elect_one()
  if (tid < 32 && p < planes) {
    if (elect_one()) {
      mbarrier_expect_tx(&full[0], bytes);
      bulk_copy_global_to_shared(buf, in + p * plane, bytes, &full[0]);
    }
    __syncwarp();
  }
  const int HW = H * W;
  uint32_t parity[2] = {0, 0};
  for (int s = 0; p < planes; p += gridDim.x, s ^= 1) {
    const long long pn = p + gridDim.x;
    // the other buffer was read in the previous trip; every thread passed the barrier at its end
    if (tid < 32 && pn < planes) {
      if (elect_one()) {
        mbarrier_expect_tx(&full[s ^ 1], bytes);
        bulk_copy_global_to_shared(buf + (size_t)(s ^ 1) * plane, in + pn * plane, bytes, &full[s ^ 1]);
      }
      __syncwarp();
    }
    mbarrier_wait(&full[s], parity[s]);
    parity[s] ^= 1;
    const float* src = buf + (size_t)s * plane;
    float* dst = out + p * HW;
    int x = tid % W, y = tid / W;
    const int dx = blockDim.x % W, dy = blockDim.x / W;
    for (int i = tid; i < HW; i += blockDim.x) {
      const ResizeAxis cx = xtab[x], cy = ytab[y];
      const float v00 = src[cy.i0 + cx.i0], v01 = src[cy.i0 + cx.i1];
      const float v10 = src[cy.i1 + cx.i0], v11 = src[cy.i1 + cx.i1];
      const float top = __fadd_rn(__fmul_rn(cx.l0, v00), __fmul_rn(cx.l1, v01));
      const float bot = __fadd_rn(__fmul_rn(cx.l0, v10), __fmul_rn(cx.l1, v11));
      dst[i] = __fadd_rn(__fmul_rn(cy.l0, top), __fmul_rn(cy.l1, bot));
      x += dx;
      y += dy;
      if (x >= W) { x -= W; ++y; }
    }
    __syncthreads();   // buffer s may be refilled two trips from now, by the copy issued in the next one
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
  cudaStream_t st = (cudaStream_t)stream;
  const long long plane = (long long)h * w;
  const size_t smem = ((16 + (size_t)(W + H) * sizeof(ResizeAxis) + 127) & ~(size_t)127) + 2 * (size_t)plane * sizeof(float);
  // staged path: planes are whole 16-byte units, two of them (plus the tables) fit a quarter of an SM's shared memory
  if (plane % 4 == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0 && smem <= 56 * 1024) {
    if (smem > 48 * 1024)
      B200BEV_CUDA_TRY(cudaFuncSetAttribute(bilinear_resize_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (int)std::min<long long>(planes, (long long)sm_count() * 4);
    bilinear_resize_staged_kernel<<<grid, 256, smem, st>>>(in, out, planes, h, w, H, W);
  } else {
    bilinear_resize_kernel<<<grid_for(planes * H * W, 256), 256, 0, st>>>(in, out, planes, h, w, H, W);
  }
  return launch_status();
}

extern "C" B200BEV_API int b200bev_bilinear_resize_nhwc_bf16(const void* in_nhwc, int B, int h, int w, int C, void* out_nhwc, int H, int W,
                                                             int C_total, int c_offset, void* stream) {
  if (!in_nhwc || !out_nhwc || B <= 0 || C <= 0 || h <= 0 || w <= 0 || H <= 0 || W <= 0) return B200BEV_ERR_INVALID_ARGUMENT;
  if (c_offset < 0 || c_offset + C > C_total) return B200BEV_ERR_INVALID_ARGUMENT;
  if ((C & 7) || (C_total & 7) || (c_offset & 7) || ((reinterpret_cast<uintptr_t>(in_nhwc) | reinterpret_cast<uintptr_t>(out_nhwc)) & 15))
    return B200BEV_ERR_UNSUPPORTED;
  const long long total = (long long)B * H * W * (C / 8);
  bilinear_resize_nhwc_bf16_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(in_nhwc), B, h, w, C, reinterpret_cast<__nv_bfloat16*>(out_nhwc), H, W, C_total, c_offset);
  return launch_status();
}
