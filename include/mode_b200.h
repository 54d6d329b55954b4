/* mode_b200.h -- C ABI of libmode_b200.so, the B200 (sm_100a) kernel library behind the
 * MODE stereo hot path (nju-ee/MODE-2022).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in `_host`;
 *   - the caller allocates every output and every workspace (reference ownership model:
 *     models/basic/spherical_conv/sphere_conv.py:35, sphere_conv_cuda.cpp:171);
 *   - `stream` is a cudaStream_t passed as void* (reference: kernels run on
 *     at::cuda::getCurrentCUDAStream(), sphere_conv_cuda_kernel.cu:280);
 *   - functions are re-entrant and keep no global state; they return 0 on success and a
 *     negative MODE_E* code otherwise; mode_b200_last_error() gives the thread-local message
 *     (reference: TORCH_CHECK/AT_ERROR -> RuntimeError, sphere_conv_cuda.cpp:43-124; the
 *     reference swallows launch errors, kernel.cu:286-289 -- this library reports them);
 *   - no function synchronises the device.
 *
 * Each entry cites the reference interface it replaces (paths relative to the reference root).
 */
#ifndef MODE_B200_H_
#define MODE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MODE_OK 0
#define MODE_EINVAL (-1)   /* bad shape / argument (reference: shape_check, sphere_conv_cuda.cpp:40-126) */
#define MODE_ECUDA (-2)    /* CUDA launch / runtime error */
#define MODE_ENOSUP (-3)   /* configuration not supported by this build */

typedef uint16_t mode_h16; /* raw 16-bit float: bfloat16 (fmt = MODE_FMT_BF16) or IEEE half (fmt = MODE_FMT_FP16) */
#define MODE_FMT_BF16 0
#define MODE_FMT_FP16 1

int mode_b200_version(void);
const char* mode_b200_last_error(void);
/* number of kernel launches issued through this library by the calling process (bench `gpu_launches`) */
unsigned long long mode_b200_launch_count(void);

/* ---- a4. concatenation cost volume ------------------------------------------------------------
 * replaces models/mode_disparity.py:104-113 (CPU zero tensor + H2D + 2*D/4 slice copies).
 *   cost[b,c,i,h,w]    = ref[b,c,h,w]   if w >= i else 0
 *   cost[b,C+c,i,h,w]  = tgt[b,c,h,w-i] if w >= i else 0          i in [0, D4)
 * f32: NCHW in, NCDHW out (bit-exact parity layout).  16: any 16-bit format, NHWC in, NDHWC out (the layout the
 * tensor-core conv3d consumes).  W % 4 == 0 (f32), C % 8 == 0 (16). */
int mode_cost_volume_f32(const float* ref, const float* tgt, float* cost, int B, int C, int H, int W, int D4, void* stream);
int mode_cost_volume_16(const mode_h16* ref, const mode_h16* tgt, mode_h16* cost, int B, int C, int H, int W, int D4, void* stream);
/* backward of the fp32 cost volume (training: autograd through the slice assignments of models/mode_disparity.py:104-113):
 * grad_cost (B,2C,D4,H,W) -> grad_ref, grad_tgt (B,C,H,W), overwritten; a deterministic gather-sum over the D4 shifts. */
int mode_cost_volume_backward_f32(const float* grad_cost, float* grad_ref, float* grad_tgt, int B, int C, int H, int W, int D4, void* stream);

/* ---- a1. stem convolution of the feature extractor (tcgen05) ------------------------------------
 * replaces sphere_feature_extraction.firstconv[0] = convbn(3, 32, 7, 2, 3, 1) + ReLU (models/submodule.py:155, 15-17, called
 * from models/mode_disparity.py:99-100): fp32 NCHW images in -- x0 (B0,3,H,W) and optionally x1 (B1,3,H,W), i.e. the
 * left and right batches without a concatenation pass -- NHWC 16-bit activation (B0+B1, Ho, Wo, 32) out, Ho = (H-1)/2+1.
 * w (32,3,7,7) fp32; y = conv*scale[co] + shift[co] (ReLU if relu); scale/shift may be NULL.  W <= 2048. */
int mode_stem_conv_tc(const float* x0, const float* x1, const float* w, const float* scale, const float* shift, mode_h16* out, int B0, int B1, int H, int W,
                      int relu, int fmt, void* stream);

/* ---- a4+a5. cost volume fused into the first 3-D convolution --------------------------------------
 * replaces the cost-volume build (models/mode_disparity.py:104-113) AND dres0[0] = convbn_3d(64, 32, 3, 1, 1) + ReLU (:66, :115)
 * without materialising the volume.  ur / ut: per-(kd,kw) column responses of the two feature maps, (B*H*W, 9, 32) fp32,
 * ur[p][kd*3+kw][o] = sum_{c,kh} W[o,c,kd,kh,kw] * ref[c, h-1+kh, w] (ut: W[o,32+c,...] and tgt) -- one library GEMM each
 * (K = 96, N = 288).  out (B, D4, H, W, 32) NDHWC 16-bit =
 * act(conv3d(cost) * scale + shift).  Exact (see costvol_conv.cu); differs from cost_volume + conv3d_tc by fp32 summation order. */
/* GEMM operand of those column responses: cols (B*H*W, 96) 16-bit, cols[(b,h,w)][kh*32 + c] = f[b, h-1+kh, w, c] (f NHWC, 32 ch). */
int mode_costvol_cols(const mode_h16* f, mode_h16* cols, int B, int H, int W, void* stream);
int mode_costvol_conv_fused(const float* ur, const float* ut, const float* scale, const float* shift, mode_h16* out, int B, int D4, int H, int W, int relu, int fmt,
                            void* stream);

/* ---- a6/a7. trilinear upsample + softmax + soft-argmin (+ confidence) --------------------------
 * replaces models/mode_disparity.py:143-152 (+131-141 for the training heads), :157-183 and
 * models/submodule.py:50-57.  cost: (B, D4, H4, W4) fp32 -> pred (B, H, W) [, conf (B, H, W) or NULL].
 * align_corners=True scales; conf = P[r] + P[clamp(r-1)] + P[clamp(r+1)], r = rint(pred). */
int mode_disp_regress(const float* cost, float* pred, float* conf, int B, int D4, int H4, int W4, int D, int H, int W, void* stream);
/* backward of one soft-argmin head (training: reference autograd through models/mode_disparity.py:131-152, three materialised
 * (B,D,H,W) volumes per head).  grad_pred (B,H,W) -> grad_cost (B,D4,H4,W4), overwritten (zero-filled by the call); the softmax
 * statistics are recomputed from `cost`, nothing else is saved by the forward.  fp32 atomics (like ATen's upsample backward). */
int mode_disp_regress_backward(const float* cost, const float* grad_pred, float* grad_cost, int B, int D4, int H4, int W4, int D, int H, int W, void* stream);
/* the same with a caller-allocated workspace of mode_disp_regress_backward_workspace_bytes() bytes (0 = this shape has no workspace path;
 * workspace may then be NULL and the call is mode_disp_regress_backward): dL/dt of every full-resolution pixel goes through the workspace
 * and the bilinear transpose is a gather -- no atomics, bit-identical from run to run. */
size_t mode_disp_regress_backward_workspace_bytes(int B, int D4, int H4, int W4, int D, int H, int W);
int mode_disp_regress_backward_ws(const float* cost, const float* grad_pred, float* grad_cost, void* workspace, int B, int D4, int H4, int W4, int D, int H, int W,
                                  void* stream);

/* ---- a2. spherical convolution forward ----------------------------------------------------------
 * replaces sphere_conv_forward_cuda (sphere_conv_cuda.cpp:129-210) = sphere_im2col_gpu_kernel
 * (sphere_conv_cuda_kernel.cu:195-262) + addmm_.  stride 1, groups 1 (the only configuration MODE
 * instantiates, models/submodule.py:128-130,161); `pos` is the (2*Kh*Kw, H, W) fp32 grid of
 * SphereConv.gen_sphere_position.
 * f32: x (B,C,H,W) NCHW, w (Co,C,Kh,Kw), bias (Co) or NULL, out (B,Co,H,W).
 * Fused epilogue (both variants): y = acc*scale[co] + shift[co] (+ residual) (ReLU if relu);
 * scale/shift/residual may be NULL. */
int mode_sphere_conv_f32(const float* x, const float* pos, const float* w, const float* scale, const float* shift,
                         const float* residual, float* out, int B, int C, int H, int W, int Co, int Kh, int Kw, int relu, void* stream);
/* ---- a3. spherical convolution backward (fp32, training) ------------------------------------------
 * replaces sphere_conv_backward_cuda (sphere_conv_cuda.cpp:213-336; Python call sphere_conv.py:57-90) = addmm_ + sphere_col2im
 * (kernel.cu:293-356) for grad_input, sphere_im2col + addmm_ for grad_weight, addmm_ with ones for grad_bias -- without
 * the two (C*Kh*Kw, H*W) column buffers.  Layouts as mode_sphere_conv_f32; grad_out (B,Co,H,W).  Like the reference the
 * three gradients are ACCUMULATED into caller-zeroed buffers (sphere_conv.py:62-64); any of grad_in / grad_w / grad_bias
 * may be NULL to skip it (x may be NULL when grad_w is, w when grad_in is). */
int mode_sphere_conv_backward_f32(const float* x, const float* pos, const float* w, const float* grad_out, float* grad_in, float* grad_w, float* grad_bias,
                                  int B, int C, int H, int W, int Co, int Kh, int Kw, void* stream);
/* the same gradients, bit-identical from run to run: every contribution is accumulated in 64-bit fixed point with integer atomics
 * (order-free) in `workspace` (mode_sphere_conv_backward_workspace_bytes() bytes, 16-byte aligned, caller-allocated, contents
 * irrelevant), then converted and ADDED into the fp32 gradients.  The reference's col2im uses fp32 atomicAdd and is not reproducible
 * (sphere_conv_cuda_kernel.cu:341-352). */
size_t mode_sphere_conv_backward_workspace_bytes(int B, int C, int H, int W, int Co, int Kh, int Kw);
int mode_sphere_conv_backward_det_f32(const float* x, const float* pos, const float* w, const float* grad_out, float* grad_in, float* grad_w, float* grad_bias,
                                      void* workspace, int B, int C, int H, int W, int Co, int Kh, int Kw, void* stream);
/* tensor-core (tcgen05) variant: x (B,H,W,C) NHWC 16-bit, w_packed from mode_sphere_conv_pack_weights,
 * out (B,H,W,Co) 16-bit, fp32 accumulation.  C % 64 == 0, Co in {64,128,192,256}, 3x3. */
/* gather table = the sampling grid pre-digested once per resolution (independent of the 16-bit format; `fmt` is accepted for ABI
 * stability): per (tap, pixel) 16 bytes -- the top-left corner's pixel index, the four bilinear weights as fp16 (0 where the
 * reference's edge rules drop the corner, kernel.cu:97-107,246) and the corner's (row, col) -- followed by the tile classes of the
 * slab kernel: per 128-pixel tile position the first line / first column / line count of its input neighbourhood, and the lists of
 * tile positions that fit the shared-memory slab ("fast") or not (polar tile columns: direct-gather kernel).
 * mode_sphere_conv_table_bytes() bytes, caller-allocated, built once by mode_sphere_conv_build_table (3 small launches). */
size_t mode_sphere_conv_table_bytes(int H, int W, int Kh, int Kw);
int mode_sphere_conv_build_table(const float* pos, void* table, int H, int W, int Kh, int Kw, int fmt, void* stream);
int mode_sphere_conv_tc(const mode_h16* x, const void* table, const mode_h16* w_packed, const float* scale, const float* shift,
                        const mode_h16* residual, mode_h16* out, int B, int C, int H, int W, int Co, int relu, int fmt, void* stream);
int mode_sphere_conv_pack_weights(const float* w /*Co,C,3,3*/, mode_h16* w_packed, int C, int Co, int fmt, void* stream);

/* ---- a5. 3-D regularisation convolutions --------------------------------------------------------
 * replace nn.Conv3d / nn.ConvTranspose3d + BatchNorm3d(eval) + residual + ReLU
 * (models/submodule.py:20-22, models/mode_disparity.py:11-46,66-80,115-129).  k=3, pad=1.
 *   mode 0: conv stride 1;  mode 1: conv stride 2;  mode 2: transposed conv stride 2, output_padding 1
 * epilogue: y = acc*scale[co] + shift[co] (+ residual) (ReLU); NULLs skip a term.
 * f32: NCDHW, w in the PyTorch layout ((Co,Ci,3,3,3); transposed: (Ci,Co,3,3,3)).  Di/Hi/Wi are INPUT dims. */
int mode_conv3d_f32(const float* x, const float* w, const float* scale, const float* shift, const float* residual, float* out,
                    int B, int Ci, int Co, int Di, int Hi, int Wi, int mode, int relu, void* stream);
/* 32 -> 1 classifier convolution (classif1/2/3[2] = nn.Conv3d(32, 1, 3, padding=1, bias=False), models/mode_disparity.py:72-80)
 * with the fp32 residual chain cost_k = classif_k(out_k) + cost_{k-1} (:126-129), as a pointwise tensor-core GEMM + shifted
 * 27-term sum: x (B,D,H,W,32) NDHWC 16-bit, w (1,32,3,3,3) fp32 (rounded to fmt inside), residual_f32 / out_f32 (B,D,H,W)
 * fp32; residual_f32 may be NULL. */
int mode_conv3d_classifier_tc(const mode_h16* x, const float* w, const float* residual_f32, float* out_f32, int B, int D, int H, int W, int fmt, void* stream);

/* tensor-core (tcgen05) variant: NDHWC 16-bit activations (fmt), weights pre-packed per tap, fp32 accumulation in TMEM.
 * out_f32 != NULL writes fp32 (B,Do,Ho,Wo,Co) instead of 16-bit (used by the 32->1 classifier, Co <= 16).
 * Ci in {32,64}; Co % 32 == 0 for 16-bit output. */
int mode_conv3d_pack_weights(const float* w, mode_h16* w_packed, int Ci, int Co, int mode, int fmt, void* stream);
int mode_conv3d_tc(const mode_h16* x, const mode_h16* w_packed, const float* scale, const float* shift, const mode_h16* residual,
                   const float* residual_f32, mode_h16* out, float* out_f32, int B, int Ci, int Co, int Di, int Hi, int Wi, int mode, int relu,
                   int fmt, void* stream);
size_t mode_conv3d_packed_weight_elems(int Ci, int Co, int mode);
/* profiling aid: when set (device pointer to >= 8*grid int64), every conv3d_bf16 launch records per-CTA {smid, start ns, end ns, items} */
int mode_conv3d_set_debug_buffer(void* dev_ptr);

/* ---- layout helpers (NCHW fp32 <-> NHWC 16-bit), used at the cuDNN / custom-kernel seams ---------- */
int mode_nchw_f32_to_nhwc_16(const float* x, mode_h16* y, int B, int C, int HW, int fmt, void* stream);
int mode_nhwc_16_to_nchw_f32(const mode_h16* x, float* y, int B, int C, int HW, int fmt, void* stream);
/* channel concatenation of three NHWC 16-bit maps (torch.cat((raw, regular, sphere), 1) in front of lastconv, models/submodule.py:198):
 * a (npix, Ca), b (npix, Cb), c (npix, Cc) -> out (npix, Ca+Cb+Cc); channel counts multiples of 8. */
int mode_concat3_nhwc_16(const mode_h16* a, const mode_h16* b, const mode_h16* c, mode_h16* out, long long npix, int Ca, int Cb, int Cc, void* stream);

/* ---- a8. disparity -> depth ----------------------------------------------------------------------
 * replaces disp2depth's triangulation, save_output_disparity_stage.py:118-133 (fp64 arithmetic on fp32 inputs,
 * as the reference executes under NumPy >= 2).  phi_l: (W) fp32 table generated on the host exactly as :118-122.
 * disp (B,H,W) fp32 -> depth (B,H,W) fp32 and/or depth64 (B,H,W) fp64 (either may be NULL). */
int mode_disp_to_depth(const float* disp, const float* phi_l, float* depth, double* depth64, int B, int H, int W, float baseline, void* stream);

/* ---- a9/a11. constant-grid bilinear resampling ---------------------------------------------------
 * replaces F.grid_sample(mode='bilinear', align_corners=True, padding_mode='border') in
 * utils/geometry.py:38,88.  src (N,C,Hs,Ws), grid (Ho,Wo,2) shared by all N, out (N,C,Ho,Wo). */
int mode_grid_sample_border(const float* src, const float* grid, float* out, int N, int C, int Hs, int Ws, int Ho, int Wo, void* stream);

/* ---- a10. z-buffer forward warp with confidence ---------------------------------------------------
 * replaces depthViewTransWithConf + __iterPixels_with_conf, utils/geometry.py:94-156 (fp64 geometry,
 * fp32 buffers, strict `<`, row-major order semantics reproduced deterministically).
 * tables: sin_phi/cos_phi (W) and sin_theta/cos_theta (H) fp32, host-generated as geometry.py:108-124.
 * Rt_host: 12 doubles on the HOST: R row-major (9) then t (3).
 * workspace: 3*H*W uint32 per map (keys), caller-allocated; depth/conf in, view2/conf2 out, all (B,H,W).
 * The source depth is read from depth64 (fp64, un-rounded output of mode_disp_to_depth) when it is not NULL,
 * else from depth (fp32); numpy's promotion rules make the two differ in the reference as well. */
int mode_depth_view_trans(const float* depth, const double* depth64, const float* conf, const float* sin_phi, const float* cos_phi, const float* sin_theta,
                          const float* cos_theta, const double* Rt_host, uint32_t* workspace, float* view2, float* conf2, int B, int H,
                          int W, void* stream);

/* ---- f1. training-mode BatchNorm2d / BatchNorm3d (batch statistics) ---------------------------------
 * replaces nn.BatchNorm2d / nn.BatchNorm3d in training mode as the reference trains them (models/submodule.py:14-30,
 * models/mode_disparity.py:11-46,66-80; train_disparity.py:147-163): y = (x - mean_B) / sqrt(var_B + eps) * gamma + beta with the
 * biased batch variance; save_mean / save_invstd (C) are returned for the backward pass, batch_var (C, unbiased, may be NULL) for a
 * caller that updates the running statistics itself, and running_mean / running_var (may be NULL) are updated in place with
 * `momentum` (unbiased variance) when given.  x, y, dy, dx: fp32, (N, C, S) contiguous (S = H*W or D*H*W), or with channels_last != 0
 * (N, S, C) in memory (torch.channels_last / channels_last_3d: what cuDNN's tensor-core conv3d kernels consume without layout
 * transforms; C a power of two in [4, 256]); gamma / beta may be NULL.
 * workspace: mode_batchnorm_workspace_bytes(C, N, S) bytes, 16-byte aligned, caller-allocated (fp64 partial sums, combined in a
 * fixed order: the statistics and parameter gradients are bit-reproducible).
 * backward: dx, dgamma (C), dbeta (C) are OVERWRITTEN (dgamma / dbeta may be NULL). */
size_t mode_batchnorm_workspace_bytes(int C, long long N, long long S, int channels_last);
int mode_batchnorm_train_fwd_f32(const float* x, const float* gamma, const float* beta, float* y, float* save_mean, float* save_invstd, float* batch_var,
                                 float* running_mean, float* running_var, void* workspace, long long N, int C, long long S, int channels_last, float eps, float momentum,
                                 void* stream);
int mode_batchnorm_train_bwd_f32(const float* x, const float* dy, const float* gamma, const float* save_mean, const float* save_invstd, float* dx, float* dgamma,
                                 float* dbeta, void* workspace, long long N, int C, long long S, int channels_last, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MODE_B200_H_ */
