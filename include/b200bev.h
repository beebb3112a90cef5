/*
 * b200bev.h — C-ABI of libb200bev.so: hand-written sm_100a CUDA kernels for the
 * BEV encode + decode hot path of meg89/bevfusion_multimodal_3d_object_detection.
 *
 * The reference has no FFI of its own (pure PyTorch); its boundary is
 * a set of Python callables.  Each entry point below names the reference callable
 * (file:line, relative to the reference checkout) whose arithmetic it replaces.  The
 * Python binding a maintainer adds is one ctypes call per entry point with
 * `tensor.data_ptr()` arguments — see INTEGRATION.md.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the comment says "host";
 *   - tensors are dense, row-major, in the layout written next to the argument;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, no
 *     entry point synchronises, allocates device memory or keeps global state
 *     (re-entrant: one thread per GPU may call concurrently);
 *   - the return value is B200BEV_OK (0) or an error code; nothing throws;
 *     codes >= B200BEV_ERR_CUDA are (B200BEV_ERR_CUDA + cudaError_t).
 */
#ifndef B200BEV_H_
#define B200BEV_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200BEV_ABI_VERSION 2

#if defined(__GNUC__)
#define B200BEV_API __attribute__((visibility("default")))
#else
#define B200BEV_API
#endif

enum b200bev_status {
  B200BEV_OK = 0,
  B200BEV_ERR_INVALID_ARGUMENT = 1, /* null pointer, non-positive size, misaligned buffer */
  B200BEV_ERR_UNSUPPORTED = 2,      /* shape outside what the kernels were built for */
  B200BEV_ERR_K_OUT_OF_RANGE = 3,   /* K > H*W: torch.topk raises "selected index k out of range"
                                       (src/centernet_target.py:432) */
  B200BEV_ERR_WORKSPACE = 4,        /* caller workspace too small */
  B200BEV_ERR_CUDA = 1000           /* + cudaError_t */
};

/* precision of the shared-MLP arithmetic (b200bev_pointnet_encode) */
enum b200bev_precision {
  B200BEV_F32 = 0,        /* fp32 FFMA, parity 1e-5 of max|ref| */
  B200BEV_BF16_TENSOR = 1 /* bf16 operands on tcgen05 tensor cores, fp32 accumulate in TMEM; parity 1e-2 */
};

/* multi-radar fusion (src/encoders.py:650-659) */
enum b200bev_radar_fusion { B200BEV_RADAR_CONCAT = 0, B200BEV_RADAR_MAX = 1, B200BEV_RADAR_MEAN = 2 };

B200BEV_API int b200bev_abi_version(void);
B200BEV_API const char* b200bev_error_string(int status);
/* SM count / compute capability of the current device; any pointer may be NULL. */
B200BEV_API int b200bev_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---------------------------------------------------------------------------------------------
 * N3 (the step in front of S1)  range filter + pad / subsample of raw LiDAR sweeps.
 * Replaces: NuScenesDataset._load_lidar_points + _pad_or_subsample, src/train_detect.py:147-189
 *   (mask = strictly inside pc_range on x, y, z; points[mask] in file order; zero rows up to max_points,
 *   or points[np.random.choice(N, max_points, replace=False)] when N >= max_points), for a whole batch.
 *   raw            (total_rows, C) f32: the rows of all frames back to back (C >= 3; columns 0..2 = x, y, z)
 *   frame_offsets  (B+1) i64, device: frame b is rows [frame_offsets[b], frame_offsets[b+1])
 *   max_frame_rows upper bound of the rows of one frame (sizes the launch and the workspace)
 *   pc_range       HOST array of 6 floats [x_min, y_min, z_min, x_max, y_max, z_max] (configs/base.yaml:48)
 *   select         optional (B, max_points) i32: output row j = the select[b,j]-th in-range point of frame b
 *                  (the caller's np.random.choice draw; an index >= count gives a zero row).  NULL: in-range
 *                  points in file order, truncated to max_points, zero rows after them.
 *   out            (B, max_points, C) f32;  count (B) i32 = in-range points of each frame (before truncation)
 *   workspace      b200bev_lidar_prepare_workspace_bytes(B, max_frame_rows, select != NULL) bytes, 16-byte aligned
 * ------------------------------------------------------------------------------------------- */
B200BEV_API size_t b200bev_lidar_prepare_workspace_bytes(int B, int64_t max_frame_rows, int with_select);
B200BEV_API int b200bev_lidar_prepare(const float* raw, const int64_t* frame_offsets, int B, int C,
                          int64_t max_frame_rows, const float* pc_range, int max_points,
                          const int32_t* select, float* out, int32_t* count,
                          void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * S1a  bin-and-sort of points by BEV cell.
 * Replaces: nothing executable in the reference; implements the cell convention of
 *   src/centernet_target.py:222-224,250-257,285 (px=(x-x_min)/voxel, int(px), flat=iy*W+ix, reject
 *   px<0 or px>=W) in fp32.
 *   points  (B,N,C) f32, column 0 = x, column 1 = y (src/train_detect.py:151-155)
 *   cell    (B,N)   i32  iy*W+ix, or -1 when the point is outside the grid (or NaN)
 *   perm    (B,N)   i32  point indices ordered by (cell, point index) — a stable counting sort;
 *                        the out-of-grid points follow, in index order
 *   offsets (B,H*W+1) i32 perm[b, offsets[b,c] : offsets[b,c+1]] are the points of cell c;
 *                        offsets[b,H*W] = number of in-grid points
 * ------------------------------------------------------------------------------------------- */
B200BEV_API int b200bev_bin_sort(const float* points, int B, int N, int C,
                     float x_min, float y_min, float voxel_x, float voxel_y, int W, int H,
                     int32_t* cell, int32_t* perm, int32_t* offsets, void* stream);

/* N3 + S1a in ONE launch ("filter + compact fused into bin-and-sort", src/train_detect.py:147-189 in front of
 * the binning): per frame one thread-block cluster filters the raw sweep by range (strict, NaN fails), compacts it in file
 * order, zero-pads to max_points — exactly b200bev_lidar_prepare without `select` (a frame with more in-range points than
 * max_points keeps the first max_points) — and then bins and sorts the rows it has just written, exactly b200bev_bin_sort
 * on them with x_min/y_min = pc_range[0..1].  Outputs of both: points (B,max_points,C), count (B), cell / perm (B,max_points),
 * offsets (B,H*W+1).  Shapes the fused kernel does not take run as two launches behind the same call. */
B200BEV_API int b200bev_lidar_prepare_bin_sort(const float* raw, const int64_t* frame_offsets, int B, int C,
                                   int64_t max_frame_rows, const float* pc_range, int max_points,
                                   float voxel_x, float voxel_y, int W, int H,
                                   float* points, int32_t* count, int32_t* cell, int32_t* perm, int32_t* offsets,
                                   void* stream);

/* ---------------------------------------------------------------------------------------------
 * S1b  fused PointNet shared-MLP + max.
 * Replaces: PointNetLiDAREncoder.forward, src/encoders.py:271-306 (eval mode, BN folded);
 *           RadarEncoder.forward, src/encoders.py:531-557.
 *   points   (B,N,C) f32
 *   params   one blob, per layer l: W_l^T (dims[l], dims[l+1]) f32 row-major, then bias_l (dims[l+1]);
 *            BatchNorm already folded in (W' = W*g/sqrt(var+eps), b' = (b-mean)*g/sqrt(var+eps)+beta)
 *   dims     host array, n_layers+1 entries: C, 64, 128, ... (n_layers <= 8, every width <= 2048)
 *   out_global (B, dims[n_layers]) or NULL: max over all N points (zero rows included) —
 *            identical to the reference's torch.max(x, 2)[0].
 *   out_canvas (B, n_cells, dims[n_layers]) channels-last or NULL: per-cell max, 0 for empty cells; needs
 *            perm/offsets from b200bev_bin_sort; points outside the grid do not reach the canvas (they
 *            still count for out_global).  With both outputs requested the MLP runs once.
 *            perm == NULL (global only) walks the points in input order.
 *   tc_params  only for B200BEV_BF16_TENSOR: blob made by b200bev_pointnet_pack_bf16 (else NULL)
 * ------------------------------------------------------------------------------------------- */
B200BEV_API int b200bev_pointnet_encode(const float* points, int B, int N, int C,
                            const float* params, const int32_t* dims, int n_layers,
                            const int32_t* perm, const int32_t* offsets, int n_cells,
                            int precision, const void* tc_params,
                            float* out_global, float* out_canvas, void* stream);

/* Bytes of the tensor-core weight image for b200bev_pointnet_pack_bf16 (0 if dims unsupported). */
B200BEV_API size_t b200bev_pointnet_pack_bf16_bytes(const int32_t* dims, int n_layers);
/* Re-tiles fp32 params (layout above, device) into the bf16 swizzled stage image the tcgen05 kernel
 * streams (device, `tc_params`). Call once per weight update. */
B200BEV_API int b200bev_pointnet_pack_bf16(const float* params, const int32_t* dims, int n_layers,
                               void* tc_params, size_t tc_bytes, void* stream);

/* fp32-accuracy tensor-core path of S1b (csrc/pointnet_mlp_split.cu): every fp32 product of the shared MLP as three fp16
 * tcgen05 products with fp32 accumulation, operands scaled by exact powers of two (weights per output channel at pack
 * time, activations per layer from the layer's own maximum, reduced on the device).  Same arithmetic contract as
 * b200bev_pointnet_encode with B200BEV_F32 — PointNetLiDAREncoder.forward, src/encoders.py:289-298, parity 1e-5 of
 * max|ref| — for the layer widths C-64-128-256-512-1024 only (pack_split_bytes returns 0 otherwise; use the FFMA kernel).
 *   image      made by b200bev_pointnet_pack_split from the same fp32 `params` blob; call once per weight update
 *   workspace  device scratch, >= b200bev_pointnet_split_workspace_bytes(1, N) (one frame per pass); with
 *              b200bev_pointnet_split_workspace_bytes(B, N) the batch runs in as few passes as ~1M points allow.
 *              Holds the fp32 activations between the five layer launches; 256-byte aligned.
 * Outputs and perm/offsets/n_cells as in b200bev_pointnet_encode. */
B200BEV_API size_t b200bev_pointnet_pack_split_bytes(const int32_t* dims, int n_layers);
B200BEV_API int b200bev_pointnet_pack_split(const float* params, const int32_t* dims, int n_layers,
                                void* image, size_t image_bytes, void* stream);
B200BEV_API size_t b200bev_pointnet_split_workspace_bytes(int B, int N);
B200BEV_API int b200bev_pointnet_encode_split(const float* points, int B, int N, int C,
                                  const int32_t* dims, int n_layers,
                                  const int32_t* perm, const int32_t* offsets, int n_cells,
                                  const void* image, float* out_global, float* out_canvas,
                                  void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * S1c  multi-radar encoder: shared MLP + max per radar, then fusion.
 * Replaces: MultiRadarEncoder.forward, src/encoders.py:628-661.
 *   radar_points host array of R device pointers, radar r is (B, n_points[r], C) f32
 *   n_points     host array, R entries (ragged radars allowed)
 *   params/dims  as above (RadarEncoder 7-32-64-128-256)
 *   fc_weight    (F, R*F) f32 row-major, fc_bias (F) — used when fusion == CONCAT, else may be NULL
 *   per_radar    workspace AND output (B,R,F) f32: the stacked per-radar maxima (src/encoders.py:647)
 *   out          (B,F) f32
 * ------------------------------------------------------------------------------------------- */
B200BEV_API int b200bev_radar_encode(const float* const* radar_points, const int32_t* n_points, int R, int B, int C,
                         const float* params, const int32_t* dims, int n_layers,
                         int fusion, const float* fc_weight, const float* fc_bias,
                         float* per_radar, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * S2  camera features -> BEV.
 * b200bev_camera_mean: camera_features.mean(dim=1), src/fusion.py:233-234.
 *   feats (B,n_cam,inner) f32 -> out (B,inner) f32, inner = C*h*w.  Sum in camera order, then
 *   divide by n_cam (same association as a sequential fp32 reduction).
 * b200bev_bilinear_resize: F.interpolate(size=(H,W), mode='bilinear', align_corners=False),
 *   src/fusion.py:242-247.  in (B,C,h,w) f32 -> out (B,C,H,W) f32.
 * b200bev_camera_project: the geometric form north_star describes (no counterpart in the reference,
 *   SURVEY §0 S2): BEV cell centres (x,y,z_plane) are taken through ego->camera extrinsics and
 *   pinhole intrinsics of each camera, the feature map is sampled bilinearly (zeros outside, the
 *   grid_sample align_corners=False convention) and averaged over the cameras that see the cell.
 *   feats      (B,n_cam,C,h,w) f32
 *   intrinsics (T,n_cam,3,3) f32, ego2cam (T,n_cam,3,4) f32 [R|t]; T is 1 (shared rig) or B
 *   img_w/img_h  pixel size the intrinsics refer to (1600x900)
 *   out        (B,C,H,W) f32
 *   uv_valid   optional (T,H*W,n_cam,3) f32 debug/table output: feature-map u, v and valid(0/1)
 *   impl       B200BEV_PROJECT_AUTO, or one of the two kernels by name (they give the same bits; the
 *              parity tests run both): _STAGED returns B200BEV_ERR_UNSUPPORTED where bands do not fit
 * ------------------------------------------------------------------------------------------- */
#define B200BEV_PROJECT_AUTO 0
#define B200BEV_PROJECT_STAGED 1
#define B200BEV_PROJECT_GATHER 2
B200BEV_API int b200bev_camera_mean(const float* feats, int B, int n_cam, int64_t inner, float* out, void* stream);
B200BEV_API int b200bev_bilinear_resize(const float* in, int B, int C, int h, int w,
                            float* out, int H, int W, void* stream);
B200BEV_API int b200bev_camera_project(const float* feats, int B, int n_cam, int C, int h, int w,
                           const float* intrinsics, const float* ego2cam, int T,
                           float img_w, float img_h,
                           float x_min, float y_min, float voxel_x, float voxel_y, float z_plane,
                           int W, int H, float* out, float* uv_valid, int impl, void* stream);

/* ---------------------------------------------------------------------------------------------
 * S3  CenterNet peak extraction and box decode.
 * b200bev_centernet_nms:  _nms, src/centernet_target.py:416-421 (= src/fusion_detection.py:784-789),
 *   kernel 3 only.  heat (B,C,H,W) f32 -> out same shape: heat * (maxpool3x3(heat) == heat).
 * b200bev_centernet_topk: _topk, src/centernet_target.py:424-452 (= src/fusion_detection.py:792-820).
 *   scores (B,C,H,W) f32 -> topk_score (B,K) f32 descending, topk_ind (B,K) i64 in [0,C*K),
 *   topk_classes (B,K) i64 (identically 0 — the reference divides an index < H*W by H*W),
 *   topk_ys/topk_xs (B,K) i64.  Ties are ordered by ascending index (torch leaves them undefined).
 * b200bev_centernet_decode: decode_centernet_predictions, src/centernet_target.py:326-413
 *   (= src/fusion_detection.py:695-781, which differs only in voxel_size) — NMS + both top-K stages
 *   + gather + box assembly in one launch, fixed-size outputs:
 *   boxes (B,K,7) f32 [x,y,z,w,l,h,yaw], scores (B,K) f32, labels (B,K) i64, velocities (B,K,2) f32,
 *   ys/xs/ind (B,K) i64 (optional, may be NULL), count (B) i32 = #rows with score > score_thresh —
 *   rows [0,count) of sample b are the reference's output rows, in the same order.
 *   workspace: b200bev_centernet_workspace_bytes(B,C,K) bytes, 16-byte aligned.
 * b200bev_centernet_decode_logits ("sigmoid fused into the NMS kernel"): the same launch fed with the
 *   heat-map head's RAW output; replaces torch.sigmoid of CenterNetHead.forward, src/fusion.py:870-871, plus the
 *   decode above.  The sigmoid is evaluated as the values are loaded (1/(1+expf(-x)), IEEE divide — the bits
 *   torch.sigmoid gives on the device) and the peak test, both top-K stages and the threshold run on its result,
 *   exactly as the reference orders them.  All other arguments as b200bev_centernet_decode.
 * ------------------------------------------------------------------------------------------- */
B200BEV_API int b200bev_centernet_nms(const float* heat, int B, int C, int H, int W, float* out, void* stream);
B200BEV_API size_t b200bev_centernet_workspace_bytes(int B, int C, int K);
B200BEV_API int b200bev_centernet_topk(const float* scores, int B, int C, int H, int W, int K,
                           float* topk_score, int64_t* topk_ind, int64_t* topk_classes,
                           int64_t* topk_ys, int64_t* topk_xs,
                           void* workspace, size_t workspace_bytes, void* stream);
B200BEV_API int b200bev_centernet_decode(const float* heatmap, const float* offset, const float* size,
                             const float* rot, const float* vel,
                             int B, int C, int H, int W, int K,
                             float voxel_size, float x_origin, float y_origin, float z_value,
                             float score_thresh,
                             float* boxes, float* scores, int64_t* labels, float* velocities,
                             int64_t* ys, int64_t* xs, int64_t* ind, int32_t* count,
                             void* workspace, size_t workspace_bytes, void* stream);
B200BEV_API int b200bev_centernet_decode_logits(const float* heatmap_logits, const float* offset, const float* size,
                                    const float* rot, const float* vel,
                                    int B, int C, int H, int W, int K,
                                    float voxel_size, float x_origin, float y_origin, float z_value,
                                    float score_thresh,
                                    float* boxes, float* scores, int64_t* labels, float* velocities,
                                    int64_t* ys, int64_t* xs, int64_t* ind, int32_t* count,
                                    void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * N2  dense layers on a small batch — the Linear stacks either side of the BEV canvas.
 * b200bev_dense_layer: out = act(x W^T + bias) in fp32 FFMA (parity 1e-5).
 *   Replaces: nn.Linear (+ nn.ReLU) of FlexibleBEVFusion.radar_proj, src/fusion.py:170-173 (applied :274), and each
 *   layer of lidar_init.
 *   x (B,K) f32; weight (O,K) f32 row-major — torch's own nn.Linear layout, no repacking; bias (O) or NULL;
 *   relu != 0 applies max(.,0); out (B,O) f32.  Large layers (K % 128 == 0, O >= 1024) take the weight-streaming
 *   kernel — each weight byte is read from HBM once per 32 batch rows — everything else one warp per output row.
 * b200bev_lidar_init: Linear(K,hidden) + ReLU + Linear(hidden,O), FlexibleBEVFusion.lidar_init,
 *   src/fusion.py:144-148 (applied :258; O = 128*25*25 = 80000, a 164 MB fp32 weight).
 *   lidar_features (B,K); w1 (hidden,K), b1 (hidden); w2 (O,hidden), b2 (O); hidden_ws (B,hidden) workspace and
 *   output (the activations between the two layers); out (B,O) — the caller views it as (B,128,25,25).
 * ------------------------------------------------------------------------------------------- */
B200BEV_API int b200bev_dense_layer(const float* x, int B, int K, const float* weight, const float* bias, int O,
                        int relu, float* out, void* stream);
B200BEV_API int b200bev_lidar_init(const float* lidar_features, int B, int K, const float* w1, const float* b1,
                       int hidden, const float* w2, const float* b2, int O,
                       float* hidden_ws, float* out, void* stream);

/* The same dense layer at fp32 accuracy ON THE TENSOR CORES (parity 1e-5): every fp32 product as three fp16 tcgen05
 * products with fp32 accumulation, from a weight image of the fp32 weight's size — for the 164 MB lidar_init.2 at batch
 * 9..64, where the FFMA form is paced by the FMA pipe instead of by the one read of the weight.
 * b200bev_dense_pack_split: weight (O,K) f32 (torch's nn.Linear layout), bias (O) or NULL -> image (once per weight
 *   update); O % 128 == 0 and K % 64 == 0, else b200bev_dense_pack_split_bytes is 0 and the calls return UNSUPPORTED.
 * b200bev_dense_layer_split: out (B,O) = act(x W^T + bias) from the image; any B (64 rows of x per launch).
 * b200bev_lidar_init_split: b200bev_lidar_init with the second layer (src/fusion.py:147) read from its image.
 */
B200BEV_API size_t b200bev_dense_pack_split_bytes(int O, int K);
B200BEV_API int b200bev_dense_pack_split(const float* weight, const float* bias, int O, int K,
                             void* image, size_t image_bytes, void* stream);
B200BEV_API int b200bev_dense_layer_split(const float* x, int B, int K, const void* image, int O,
                              int relu, float* out, void* stream);
B200BEV_API int b200bev_lidar_init_split(const float* lidar_features, int B, int K, const float* w1, const float* b1,
                             int hidden, const void* image2, int O,
                             float* hidden_ws, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * N1  the convolution blocks between the hot-path kernels, bf16 on tcgen05 tensor cores (parity 1e-2).
 * b200bev_conv_bn_relu_bf16: Conv2d(k=3, padding=1 | k=1) [+ BatchNorm2d eval, folded by the caller] [+ ReLU] as one
 *   implicit-GEMM launch.  Replaces the blocks of FlexibleBEVFusion — camera_proj src/fusion.py:126-133, lidar_upsample
 *   :151-166, radar_refine :176-183, bev_fusion :199-207 — and CenterNetHead's convolutions src/fusion.py:822-854.
 *   x_nhwc       (B,H,W,Cin) bf16 channels-last, Cin a multiple of 64 (b200bev_nchw_to_nhwc_bf16 makes it)
 *   weight_image made by b200bev_conv_pack_bf16 from the folded (Cout,Cin,kh,kw) fp32 weight; taps = kh*kw = 9 or 1
 *   bias         (Cout) f32 folded bias or NULL;  relu != 0 applies max(.,0)
 *   out_nchw     (B,Cout,H,W) f32 — the layout the reference's next module takes
 * b200bev_conv_pack_bf16: re-tiles a weight into the 16 KB swizzled stages the kernel streams (once per weight update);
 *   b200bev_conv_pack_bytes gives the image size (0: unsupported shape).
 * b200bev_nchw_to_nhwc_bf16: (B,C,H,W) f32 -> channels [c_offset, c_offset+C) of a (B,H,W,C_total) bf16 tensor; writing
 *   the parts of a concatenated input into their slices replaces torch.cat of src/fusion.py:292.
 * b200bev_bilinear_resize_nhwc_bf16: F.interpolate(size=(H,W), mode='bilinear', align_corners=False) of src/fusion.py:242-247
 *   between two convolutions of the bf16 path: in (B,h,w,C) bf16 channels-last -> channels [c_offset, c_offset+C) of
 *   (B,H,W,C_total) bf16; fp32 interpolation in b200bev_bilinear_resize's op order, rounded to bf16.  C, C_total, c_offset % 8 == 0.
 * b200bev_camera_mean_nhwc_bf16: camera_features.mean(dim=1) (src/fusion.py:233-234) fused with that layout step:
 *   feats (B,n_cam,C,H,W) f32 -> channels [c_offset, c_offset+C) of (B,H,W,C_total) bf16 = the bf16 rounding of
 *   b200bev_camera_mean's result (same summation order, IEEE divide).  Needs H*W % 4 == 0 and C_total, c_offset % 8 == 0.
 * ------------------------------------------------------------------------------------------- */
B200BEV_API size_t b200bev_conv_pack_bytes(int Cout, int Cin, int taps);
B200BEV_API int b200bev_camera_mean_nhwc_bf16(const float* feats, int B, int n_cam, int C, int H, int W,
                                  void* out_nhwc, int C_total, int c_offset, void* stream);
B200BEV_API int b200bev_conv_pack_bf16(const float* weight, int Cout, int Cin, int taps,
                           void* image, size_t image_bytes, void* stream);
B200BEV_API int b200bev_nchw_to_nhwc_bf16(const float* in, int B, int C, int H, int W,
                              void* out_nhwc, int C_total, int c_offset, void* stream);
B200BEV_API int b200bev_bilinear_resize_nhwc_bf16(const void* in_nhwc, int B, int h, int w, int C,
                                      void* out_nhwc, int H, int W, int C_total, int c_offset, void* stream);
B200BEV_API int b200bev_conv_bn_relu_bf16(const void* x_nhwc, int B, int H, int W, int Cin,
                              const void* weight_image, const float* bias, int Cout, int taps, int relu,
                              float* out_nchw, void* stream);
/* Same launch writing the result as the NEXT convolution's input: channels [out_c_offset, out_c_offset + Cout) of a
 * (B,H,W,out_c_total) bf16 channels-last tensor (a conv -> conv chain, or a branch writing its slice of the concatenated
 * bev_fusion input, src/fusion.py:292) — no fp32 round trip and no layout pass in between.  out_nchw may be given as
 * well (both are written) or NULL. */
B200BEV_API int b200bev_conv_bn_relu_bf16_nhwc(const void* x_nhwc, int B, int H, int W, int Cin,
                                   const void* weight_image, const float* bias, int Cout, int taps, int relu,
                                   void* out_nhwc, int out_c_total, int out_c_offset,
                                   float* out_nchw, void* stream);

/* The same convolution blocks at FP32 ACCURACY (parity 1e-5 of max|ref|, the default precision of the drop-in): every fp32
 * product as three fp16 tcgen05 products (w_hi.x_hi + w_hi.x_lo + w_lo.x_hi, fp32 accumulation), operands scaled by exact
 * powers of two — weights per output channel at pack time, the input tensor by its own maximum, reduced on the device.
 * Replaces the reference's fp32 cuDNN layers (src/fusion.py:126-133,151-166,176-183,199-207,822-854) without TF32.
 *   b200bev_absmax               atomicMax of max|x| (as float bits) into *stat (4 bytes, zeroed by the caller); call it
 *                                once per input part before the layout kernel
 *   b200bev_nchw_to_nhwc_split   (B,C,H,W) f32 -> channels [c_offset, c_offset+C) of BOTH halves of a
 *                                (B,H,W,[hi C_total | lo C_total]) fp16 tensor, scaled by 2^k(*stat)
 *   b200bev_conv_pack_split      (Cout,Cin,kh,kw) f32 -> image (stage triples hi,hi,lo per 64-channel chunk + per-channel
 *                                unscale); _bytes gives its size (0: unsupported)
 *   b200bev_conv_bn_relu_split   x_split + x_stat + image + bias -> (B,Cout,H,W) f32 */
B200BEV_API int b200bev_absmax(const float* x, int64_t n, void* stat, void* stream);
B200BEV_API int b200bev_nchw_to_nhwc_split(const float* in, int B, int C, int H, int W, void* out_nhwc,
                               int C_total, int c_offset, const void* stat, void* stream);
B200BEV_API size_t b200bev_conv_pack_split_bytes(int Cout, int Cin, int taps);
B200BEV_API int b200bev_conv_pack_split(const float* weight, int Cout, int Cin, int taps,
                            void* image, size_t image_bytes, void* stream);
B200BEV_API int b200bev_conv_bn_relu_split(const void* x_split, const void* x_stat, int B, int H, int W, int Cin,
                               const void* weight_image, const float* bias, int Cout, int taps, int relu,
                               float* out_nchw, void* stream);

/* The radar branch of FlexibleBEVFusion.forward (src/fusion.py:274-281) feeds `radar_refine` a spatially CONSTANT image:
 * the (B,C) projection broadcast to (B,C,H,W).  k 3x3 convolutions (padding 1) of a constant image have (2k+1)^2
 * distinct output pixels per channel, so the stack runs on an s x s image (s = 2k+1) and this entry point spreads it:
 * out[y][x] = small[cls(y)][cls(x)], cls(i) = i if i < k, s - (n - i) if i >= n - k, else k.  Same values as the full-size
 * stack (same kernels, same operands per pixel), 1/100 of the work at 50x50.
 *   small (B,C,s,s) f32; out_nchw (B,C,H,W) f32 or NULL; out_nhwc (B,H,W,C_total) bf16 or NULL, channels from c_offset. */
B200BEV_API int b200bev_border_expand(const float* small, int B, int C, int s, int H, int W,
                          float* out_nchw, void* out_nhwc, int C_total, int c_offset, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200BEV_H_ */
