// disp_regress.cu -- fused trilinear-upsample (align_corners) + softmax over D + soft-argmin (+ confidence).
//
// Reference: models/mode_disparity.py:143-152 (F.upsample trilinear -> softmax -> disparityregression,
// models/submodule.py:50-57) and :157-183 (confidence = P[r] + P[r-1] + P[r+1], border clamp).
// The reference materialises two (B,D,H,W) fp32 volumes (403 MB each at 1024x512/D=192); this kernel reads the
// 1/4-resolution logits (6.3 MB, L2 resident) and writes only the two (B,H,W) maps.
//
// Arithmetic follows ATen's upsample_trilinear3d: src = scale*dst with scale=(in-1)/(out-1) in fp32,
// i0=(int)src, i1=i0+(i0<in-1), l1=src-i0, l0=1-l1, value = ld0*(lh0*(lw0*a+lw1*b)+lh1*(lw0*c+lw1*d)) + ld1*(...).
// The inner (h,w) bilinear is depth independent, so it is evaluated once per 1/4-res depth plane (D/4 values,
// kept in shared memory as t[d4][thread]) and the d-lerp + exp runs over the D fine planes from there.
// One thread per output pixel; consecutive threads = consecutive w (coalesced output, ~8 distinct low-res
// columns per warp so logit loads are L1 broadcasts).
#include "common.cuh"
using namespace mode;

constexpr int kRegThreads = 256;

// exp(y) for y <= 0 as MUFU.EX2 of y*log2(e), flush-to-zero: the same bits as __expf except that results below 2^-126
// (y < -87.3, i.e. weights that cannot change an fp32 sum that is >= 1) become 0 instead of a denormal; saves the range
// fix-up (2 FMUL + FSETP per call) in a kernel that is issue bound
__device__ __forceinline__ float exp_neg(float y) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(y * 1.4426950408889634f));
  return r;
}

// D4C / DC: compile-time depth sizes (48 / 192 = the model's maxdisp 192) or 0 / 0 = run-time sizes.  With constants the
// fine-depth loop unrolls completely: every plane's (d0, d1, l0, l1) folds into immediates (same fp32 expressions, evaluated
// by the compiler) and the t[] reads of neighbouring planes merge -- 7 instead of ~15 instructions per plane and pixel in
// a kernel that is issue bound (604 M exponentials per 6 pairs, ~3.5 k instructions per pixel before).
template <int D4C, int DC>
__global__ void __launch_bounds__(kRegThreads) disp_regress_kernel(const float* __restrict__ cost, float* __restrict__ pred,
                                                                   float* __restrict__ conf, int D4r, int H4, int W4, int Dr, int H, int W,
                                                                   float sdr, float sh, float sw) {
  const int D4 = D4C ? D4C : D4r, D = DC ? DC : Dr;
  const float sd = DC ? (float)(D4C - 1) / (float)(DC > 1 ? DC - 1 : 1) : sdr;
  extern __shared__ float t_s[];  // [D4][kRegThreads]
  const int b = blockIdx.y;
  const int pix = blockIdx.x * kRegThreads + threadIdx.x;
  const bool active = pix < H * W;
  const int p = active ? pix : H * W - 1;
  const int h = p / W, w = p - h * W;
  const float hs = sh * h, ws = sw * w;
  const int h0 = (int)hs, w0 = (int)ws;
  const int h1 = h0 + (h0 < H4 - 1), w1 = w0 + (w0 < W4 - 1);
  const float lh1 = hs - h0, lw1 = ws - w0;
  const float lh0 = 1.f - lh1, lw0 = 1.f - lw1;
  const float* cb = cost + (size_t)b * D4 * H4 * W4;
  const int o00 = h0 * W4 + w0, o01 = h0 * W4 + w1, o10 = h1 * W4 + w0, o11 = h1 * W4 + w1;
  float* t = t_s + threadIdx.x;
  float m = -INFINITY;
#pragma unroll 4
  for (int d4 = 0; d4 < D4; ++d4) {
    const float* cp = cb + (size_t)d4 * H4 * W4;
    float v = lh0 * (lw0 * __ldg(cp + o00) + lw1 * __ldg(cp + o01)) + lh1 * (lw0 * __ldg(cp + o10) + lw1 * __ldg(cp + o11));
    t[d4 * kRegThreads] = v;
    m = fmaxf(m, v);
  }
  float sum = 0.f, wsum = 0.f;
  auto plane = [&](int d) {
    const float ds = sd * d;
    const int d0 = (int)ds;
    const int d1 = d0 + (d0 < D4 - 1);
    const float l1 = ds - d0, l0 = 1.f - l1;
    const float x = l0 * t[d0 * kRegThreads] + l1 * t[d1 * kRegThreads];
    const float e = exp_neg(x - m);
    sum += e;
    wsum = fmaf(e, (float)d, wsum);
  };
  if (DC > 0) {
#pragma unroll
    for (int d = 0; d < (DC > 0 ? DC : 1); ++d) plane(d);
  } else {
#pragma unroll 4
    for (int d = 0; d < D; ++d) plane(d);
  }
  const float pr = wsum / sum;
  if (!active) return;
  pred[(size_t)b * H * W + pix] = pr;
  if (conf != nullptr) {
    const int r = (int)rintf(pr);  // torch.round: half-to-even (mode_disparity.py:159)
    const float inv = 1.f / sum;
    float c = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {  // order of the reference's three grid_samples: r, r-1, r+1
      int d = r + (k == 0 ? 0 : (k == 1 ? -1 : 1));
      d = min(max(d, 0), D - 1);
      const float ds = sd * d;
      const int d0 = (int)ds;
      const int d1 = d0 + (d0 < D4 - 1);
      const float l1 = ds - d0, l0 = 1.f - l1;
      const float x = l0 * t[d0 * kRegThreads] + l1 * t[d1 * kRegThreads];
      c += exp_neg(x - m) * inv;
    }
    conf[(size_t)b * H * W + pix] = c;
  }
}

// ---- maxdisp 192 at image widths that are multiples of 256 (every BASELINE config): the kernel above is instruction-issue bound
// (ncu r02: 2700 warp instructions per 32 pixels against 1536 MUFU cycles, issue slots 54 % used, long-scoreboard stalls on the
// 192 L1/L2 logit loads per thread).  Here
//   * a block = 256 consecutive pixels of ONE image row, so the row interpolation weights (lh0, lh1) are block constants: the block
//     first builds the row-interpolated coarse logits u[d4][col] = lh0 * c[d4][h0][col] + lh1 * c[d4][h1][col] for its <= 66 coarse
//     columns in shared memory (coalesced loads, all in flight at once), and a pixel's 48 bilinear logits are
//     t[d4] = lw0 * u[d4][w0] + lw1 * u[d4][w1] -- ATen's trilinear formula with the h- and w-lerp exchanged (same value up to one
//     fp32 rounding of a logit, far inside the 1e-4 disparity budget; measured in tests/test_gpu_kernels.py);
//   * those 48 logits stay in REGISTERS (every index of the unrolled fine-depth loop is a compile-time constant), pre-shifted by the
//     maximum and pre-scaled by log2(e): a fine plane is FFMA (depth lerp t0 + l1 (t1 - t0), the difference shared by the ~4 planes
//     of a coarse interval), MUFU.EX2, FADD, FFMA -- 4.25 instructions against 8 MUFU cycles per warp;
//   * the three confidence planes (run-time indices) re-evaluate their logits from the shared-memory tile with the same
//     expressions, so P[r], P[r-1], P[r+1] are the very numbers the softmax sum was built from.
constexpr int kTileCols = 66;  // 255 * (W4-1)/(W-1) < 63.75 -> at most 64 + 2 coarse columns per block

__global__ void __launch_bounds__(kRegThreads, 3) disp_regress_192_kernel(const float* __restrict__ cost, float* __restrict__ pred, float* __restrict__ conf, int H4,
                                                                          int W4, int H, int W, float sh, float sw) {
  constexpr int D4 = 48, D = 192;
  constexpr float sd = (float)(D4 - 1) / (float)(D - 1);
  constexpr float kLog2e = 1.4426950408889634f;
  __shared__ float tile[D4][kTileCols];
  const int wblks = W / kRegThreads;
  const int h = blockIdx.x / wblks, wb = (blockIdx.x - h * wblks) * kRegThreads;
  const int b = blockIdx.y;
  const float hs = sh * h;
  const int h0 = (int)hs;
  const int h1 = h0 + (h0 < H4 - 1);
  const float lh1 = hs - h0, lh0 = 1.f - lh1;
  const int c_lo = (int)(sw * wb);  // w0 of the block's first pixel (w0 is monotone in w)
  {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* r0 = cost + ((size_t)b * D4 * H4 + h0) * W4;
    const float* r1 = cost + ((size_t)b * D4 * H4 + h1) * W4;
    const int g0 = min(c_lo + lane, W4 - 1), g1 = min(c_lo + lane + 32, W4 - 1), g2 = min(c_lo + lane + 64, W4 - 1);
#pragma unroll
    for (int j = 0; j < D4 / 8; ++j) {
      const int d4 = warp + 8 * j;
      const size_t po = (size_t)d4 * H4 * W4;
      tile[d4][lane] = lh0 * __ldg(r0 + po + g0) + lh1 * __ldg(r1 + po + g0);
      tile[d4][lane + 32] = lh0 * __ldg(r0 + po + g1) + lh1 * __ldg(r1 + po + g1);
      if (lane < kTileCols - 64) tile[d4][lane + 64] = lh0 * __ldg(r0 + po + g2) + lh1 * __ldg(r1 + po + g2);
    }
  }
  __syncthreads();
  const int w = wb + threadIdx.x;
  const float ws = sw * w;
  const int w0 = (int)ws;
  const int w1 = w0 + (w0 < W4 - 1);
  const float lw1 = ws - w0, lw0 = 1.f - lw1;
  const int i0 = w0 - c_lo, i1 = w1 - c_lo;
  auto coarse = [&](int d4) { return lw0 * tile[d4][i0] + lw1 * tile[d4][i1]; };
  float t[D4];
  float m = -INFINITY;
#pragma unroll
  for (int d4 = 0; d4 < D4; ++d4) {
    t[d4] = coarse(d4);
    m = fmaxf(m, t[d4]);
  }
  const float mneg = -m * kLog2e;
#pragma unroll
  for (int d4 = 0; d4 < D4; ++d4) t[d4] = fmaf(t[d4], kLog2e, mneg);  // (t - m) log2(e)
  float sum = 0.f, wsum = 0.f;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    const float ds = sd * d;
    const int d0 = (int)ds;
    const int d1 = d0 + (d0 < D4 - 1);
    const float l1 = ds - d0;
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(l1, t[d1] - t[d0], t[d0])));
    sum += e;
    wsum = fmaf(e, (float)d, wsum);
  }
  const float pr = wsum / sum;
  const size_t o = ((size_t)b * H + h) * W + w;
  pred[o] = pr;
  if (conf != nullptr) {
    const int r = (int)rintf(pr);  // torch.round: half-to-even (mode_disparity.py:159)
    const float inv = 1.f / sum;
    float c = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {  // order of the reference's three grid_samples: r, r-1, r+1
      int d = r + (k == 0 ? 0 : (k == 1 ? -1 : 1));
      d = min(max(d, 0), D - 1);
      const float ds = sd * d;
      const int d0 = (int)ds;
      const int d1 = d0 + (d0 < D4 - 1);
      const float l1 = ds - d0;
      const float t0 = fmaf(coarse(d0), kLog2e, mneg), t1 = fmaf(coarse(d1), kLog2e, mneg);
      float e;
      asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(l1, t1 - t0, t0)));
      c += e * inv;
    }
    conf[o] = c;
  }
}

extern "C" int mode_disp_regress(const float* cost, float* pred, float* conf, int B, int D4, int H4, int W4, int D, int H, int W,
                                 void* stream) {
  MODE_CHECK_ARG(cost && pred, "disp_regress: null pointer");
  MODE_CHECK_ARG(B > 0 && D4 > 0 && H4 > 0 && W4 > 0 && D > 1 && H > 1 && W > 1, "disp_regress: bad shape");
  MODE_CHECK_ARG(D4 <= 128, "disp_regress: D/4 = %d > 128 not supported", D4);
  const float sd = (float)(D4 - 1) / (float)(D - 1), sh = (float)(H4 - 1) / (float)(H - 1), sw = (float)(W4 - 1) / (float)(W - 1);
  const size_t smem = (size_t)D4 * kRegThreads * sizeof(float);
  static thread_local int attr_dev[kMaxDevices] = {};  // the attribute is per device and per function
  int& attr_set_for = attr_dev[current_device()];
  if (smem > 48 * 1024 && attr_set_for < (int)smem) {
    MODE_CHECK_CUDA(cudaFuncSetAttribute(disp_regress_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "disp_regress");
    MODE_CHECK_CUDA(cudaFuncSetAttribute(disp_regress_kernel<48, 192>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "disp_regress");
    attr_set_for = (int)smem;
  }
  if (D4 == 48 && D == 192 && W % kRegThreads == 0 && W4 * 4 == W && W4 >= 2) {
    dim3 grid192((unsigned)((long long)H * (W / kRegThreads)), B);
    disp_regress_192_kernel<<<grid192, kRegThreads, 0, (cudaStream_t)stream>>>(cost, pred, conf, H4, W4, H, W, sh, sw);
    MODE_CHECK_LAUNCH("disp_regress");
    return MODE_OK;
  }
  dim3 grid(ceil_div((long long)H * W, kRegThreads), B);
  if (D4 == 48 && D == 192)
    disp_regress_kernel<48, 192><<<grid, kRegThreads, smem, (cudaStream_t)stream>>>(cost, pred, conf, D4, H4, W4, D, H, W, sd, sh, sw);
  else
    disp_regress_kernel<0, 0><<<grid, kRegThreads, smem, (cudaStream_t)stream>>>(cost, pred, conf, D4, H4, W4, D, H, W, sd, sh, sw);
  MODE_CHECK_LAUNCH("disp_regress");
  return MODE_OK;
}


// ---- backward (training): d loss / d logits of one soft-argmin head, fused -- the reference back-propagates through three
// materialised (B,D,H,W) volumes per head (upsample, softmax, the d-weighted sum: models/mode_disparity.py:131-152).  Here a
// thread owns one full-resolution pixel: it re-evaluates the pixel's D/4 bilinear logits t[], the softmax statistics and the
// prediction exactly as the forward kernel does, then
//     d L / d u_d   = g * p_d * (d - pred)                      (soft-argmin through the softmax)
//     d L / d t[a] += l_a(d) * dL/du_d                           (transpose of the depth lerp; per-thread column in shared memory)
//     d L / d cost[a, h_j, w_k] += lh_j * lw_k * dL/dt[a]        (transpose of the (h,w) bilinear)
// The last step first accumulates into a shared-memory tile of the coarse volume covering the block's pixels (a block = up to 256
// consecutive pixels of ONE image row: 2 coarse rows x <= 66 columns x D/4), then adds the tile to global memory -- 8x fewer global
// atomics than one per (pixel, plane, corner).  fp32 atomics: summation order varies from run to run, like ATen's own
// upsample_trilinear3d_backward.
constexpr int kBwdCols = 72;  // coarse columns a block of 256 pixels can touch (256 * (W4-1)/(W-1) + 2 <= 66) + slack

__global__ void __launch_bounds__(kRegThreads) disp_regress_bwd_kernel(const float* __restrict__ cost, const float* __restrict__ gpred, float* __restrict__ gcost, int D4,
                                                                       int H4, int W4, int D, int H, int W, float sd, float sh, float sw) {
  extern __shared__ float bw_s[];
  float* t_s = bw_s;                                   // [D4][kRegThreads] bilinear logits
  float* gt_s = bw_s + (size_t)D4 * kRegThreads;       // [D4][kRegThreads] gradient w.r.t. t
  float* acc = gt_s + (size_t)D4 * kRegThreads;        // [D4][2][kBwdCols] coarse tile
  const int b = blockIdx.z, h = blockIdx.y;
  const int w = blockIdx.x * kRegThreads + threadIdx.x;
  const bool active = w < W;
  const int wc = active ? w : W - 1;
  const float hs = sh * h, ws = sw * wc;
  const int h0 = (int)hs, w0 = (int)ws;
  const int h1 = h0 + (h0 < H4 - 1), w1 = w0 + (w0 < W4 - 1);
  const float lh1 = hs - h0, lw1 = ws - w0;
  const float lh0 = 1.f - lh1, lw0 = 1.f - lw1;
  const int wbase = (int)(sw * (blockIdx.x * kRegThreads));  // first coarse column of this block (w0 is monotone in w)
  for (int i = threadIdx.x; i < D4 * 2 * kBwdCols; i += kRegThreads) acc[i] = 0.f;
  const float* cb = cost + (size_t)b * D4 * H4 * W4;
  const int o00 = h0 * W4 + w0, o01 = h0 * W4 + w1, o10 = h1 * W4 + w0, o11 = h1 * W4 + w1;
  float* t = t_s + threadIdx.x;
  float* gt = gt_s + threadIdx.x;
  float m = -INFINITY;
  for (int d4 = 0; d4 < D4; ++d4) {
    const float* cp = cb + (size_t)d4 * H4 * W4;
    const float v = lh0 * (lw0 * __ldg(cp + o00) + lw1 * __ldg(cp + o01)) + lh1 * (lw0 * __ldg(cp + o10) + lw1 * __ldg(cp + o11));
    t[d4 * kRegThreads] = v;
    gt[d4 * kRegThreads] = 0.f;
    m = fmaxf(m, v);
  }
  float sum = 0.f, wsum = 0.f;
  for (int d = 0; d < D; ++d) {
    const float ds = sd * d;
    const int d0 = (int)ds;
    const int d1 = d0 + (d0 < D4 - 1);
    const float l1 = ds - d0, l0 = 1.f - l1;
    const float e = exp_neg(l0 * t[d0 * kRegThreads] + l1 * t[d1 * kRegThreads] - m);
    sum += e;
    wsum = fmaf(e, (float)d, wsum);
  }
  const float pr = wsum / sum;
  const float g = active ? __ldg(gpred + ((size_t)b * H + h) * W + wc) / sum : 0.f;
  for (int d = 0; d < D; ++d) {
    const float ds = sd * d;
    const int d0 = (int)ds;
    const int d1 = d0 + (d0 < D4 - 1);
    const float l1 = ds - d0, l0 = 1.f - l1;
    const float e = exp_neg(l0 * t[d0 * kRegThreads] + l1 * t[d1 * kRegThreads] - m);
    const float gu = g * e * ((float)d - pr);
    gt[d0 * kRegThreads] += l0 * gu;
    gt[d1 * kRegThreads] += l1 * gu;
  }
  __syncthreads();  // acc is zeroed
  if (active) {
    const int c0 = w0 - wbase, c1 = w1 - wbase;
    const float k00 = lh0 * lw0, k01 = lh0 * lw1, k10 = lh1 * lw0, k11 = lh1 * lw1;
    for (int d4 = 0; d4 < D4; ++d4) {
      const float v = gt[d4 * kRegThreads];
      float* a = acc + (size_t)d4 * 2 * kBwdCols;
      atomicAdd(a + c0, k00 * v);
      atomicAdd(a + c1, k01 * v);
      atomicAdd(a + kBwdCols + c0, k10 * v);
      atomicAdd(a + kBwdCols + c1, k11 * v);
    }
  }
  __syncthreads();
  // the tile's two coarse rows are h0 and h1 (uniform over the block: one image row per block); h1 == h0 at the bottom edge
  float* gb = gcost + (size_t)b * D4 * H4 * W4;
  const int ncol = min(kBwdCols, W4 - wbase);
  for (int i = threadIdx.x; i < D4 * 2 * ncol; i += kRegThreads) {
    const int c = i % ncol, r = (i / ncol) & 1, d4 = i / (2 * ncol);
    const float v = acc[((size_t)d4 * 2 + r) * kBwdCols + c];
    if (v != 0.f) atomicAdd(gb + ((size_t)d4 * H4 + (r ? h1 : h0)) * W4 + wbase + c, v);
  }
}

extern "C" int mode_disp_regress_backward(const float* cost, const float* grad_pred, float* grad_cost, int B, int D4, int H4, int W4, int D, int H, int W, void* stream) {
  MODE_CHECK_ARG(cost && grad_pred && grad_cost, "disp_regress_backward: null pointer");
  MODE_CHECK_ARG(B > 0 && D4 > 0 && H4 > 0 && W4 > 0 && D > 1 && H > 1 && W > 1, "disp_regress_backward: bad shape");
  MODE_CHECK_ARG(D4 <= 64, "disp_regress_backward: D/4 = %d > 64 not supported", D4);
  MODE_CHECK_ARG((long long)kRegThreads * (W4 - 1) / (W - 1) + 3 <= kBwdCols, "disp_regress_backward: upsampling factor W/W4 too small for the coarse tile");
  const float sd = (float)(D4 - 1) / (float)(D - 1), sh = (float)(H4 - 1) / (float)(H - 1), sw = (float)(W4 - 1) / (float)(W - 1);
  const size_t smem = ((size_t)2 * D4 * kRegThreads + (size_t)D4 * 2 * kBwdCols) * sizeof(float);
  static thread_local size_t attr_dev[kMaxDevices] = {};
  size_t& attr = attr_dev[current_device()];
  if (smem > 48 * 1024 && smem > attr) {
    MODE_CHECK_CUDA(cudaFuncSetAttribute(disp_regress_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "disp_regress_backward");
    attr = smem;
  }
  MODE_CHECK_CUDA(cudaMemsetAsync(grad_cost, 0, (size_t)B * D4 * H4 * W4 * sizeof(float), (cudaStream_t)stream), "disp_regress_backward");
  dim3 grid(ceil_div(W, kRegThreads), H, B);
  disp_regress_bwd_kernel<<<grid, kRegThreads, smem, (cudaStream_t)stream>>>(cost, grad_pred, grad_cost, D4, H4, W4, D, H, W, sd, sh, sw);
  MODE_CHECK_LAUNCH("disp_regress_backward");
  return MODE_OK;
}

// ---- backward, maxdisp 192 / width % 256 == 0: two deterministic passes instead of shared + global fp32 atomics.
//   pass 1 (row-tile kernel, as the forward): a pixel's 48 coarse logits and their gradients live in registers (compile-time indices);
//           dL/dt[48] of every full-resolution pixel goes to a (B, 48, H, W) workspace, coalesced along w;
//   pass 2: the transpose of the (h, w) bilinear as a GATHER -- a block owns one coarse row of one plane, first folds the ~11 fine rows
//           that touch it into a row buffer (weights lh0 / lh1 of the rows whose h0 / h1 is this coarse row), then every coarse column sums
//           its <= 11 fine columns.  Fixed summation order: the gradient is bit-identical from run to run (the atomics version is not).
__global__ void __launch_bounds__(kRegThreads, 1) disp_regress_bwd192_pix_kernel(const float* __restrict__ cost, const float* __restrict__ gpred, float* __restrict__ gt_ws,
                                                                                 int H4, int W4, int H, int W, float sh, float sw) {
  constexpr int D4 = 48, D = 192;
  constexpr float sd = (float)(D4 - 1) / (float)(D - 1);
  constexpr float kLog2e = 1.4426950408889634f;
  __shared__ float tile[D4][kTileCols];
  const int wblks = W / kRegThreads;
  const int h = blockIdx.x / wblks, wb = (blockIdx.x - h * wblks) * kRegThreads;
  const int b = blockIdx.y;
  const float hs = sh * h;
  const int h0 = (int)hs;
  const int h1 = h0 + (h0 < H4 - 1);
  const float lh1 = hs - h0, lh0 = 1.f - lh1;
  const int c_lo = (int)(sw * wb);
  {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* r0 = cost + ((size_t)b * D4 * H4 + h0) * W4;
    const float* r1 = cost + ((size_t)b * D4 * H4 + h1) * W4;
    const int g0 = min(c_lo + lane, W4 - 1), g1 = min(c_lo + lane + 32, W4 - 1), g2 = min(c_lo + lane + 64, W4 - 1);
#pragma unroll
    for (int j = 0; j < D4 / 8; ++j) {
      const int d4 = warp + 8 * j;
      const size_t po = (size_t)d4 * H4 * W4;
      tile[d4][lane] = lh0 * __ldg(r0 + po + g0) + lh1 * __ldg(r1 + po + g0);
      tile[d4][lane + 32] = lh0 * __ldg(r0 + po + g1) + lh1 * __ldg(r1 + po + g1);
      if (lane < kTileCols - 64) tile[d4][lane + 64] = lh0 * __ldg(r0 + po + g2) + lh1 * __ldg(r1 + po + g2);
    }
  }
  __syncthreads();
  const int w = wb + threadIdx.x;
  const float ws = sw * w;
  const int w0 = (int)ws;
  const int w1 = w0 + (w0 < W4 - 1);
  const float lw1 = ws - w0, lw0 = 1.f - lw1;
  const int i0 = w0 - c_lo, i1 = w1 - c_lo;
  float t[D4], gt[D4];
  float m = -INFINITY;
#pragma unroll
  for (int d4 = 0; d4 < D4; ++d4) {
    t[d4] = lw0 * tile[d4][i0] + lw1 * tile[d4][i1];
    m = fmaxf(m, t[d4]);
    gt[d4] = 0.f;
  }
  const float mneg = -m * kLog2e;
#pragma unroll
  for (int d4 = 0; d4 < D4; ++d4) t[d4] = fmaf(t[d4], kLog2e, mneg);
  float sum = 0.f, wsum = 0.f;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    const float ds = sd * d;
    const int d0 = (int)ds;
    const int d1 = d0 + (d0 < D4 - 1);
    const float l1 = ds - d0;
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(l1, t[d1] - t[d0], t[d0])));
    sum += e;
    wsum = fmaf(e, (float)d, wsum);
  }
  const float pr = wsum / sum;
  const float g = __ldg(gpred + ((size_t)b * H + h) * W + w) / sum;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    const float ds = sd * d;
    const int d0 = (int)ds;
    const int d1 = d0 + (d0 < D4 - 1);
    const float l1 = ds - d0, l0 = 1.f - l1;
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(l1, t[d1] - t[d0], t[d0])));
    const float gu = g * e * ((float)d - pr);  // dL/du_d = g p_d (d - pred)
    gt[d0] = fmaf(l0, gu, gt[d0]);
    gt[d1] = fmaf(l1, gu, gt[d1]);
  }
  float* o = gt_ws + (((size_t)b * D4) * H + h) * W + w;
#pragma unroll
  for (int d4 = 0; d4 < D4; ++d4) o[(size_t)d4 * H * W] = gt[d4];
}

__global__ void __launch_bounds__(kRegThreads) disp_regress_bwd192_gather_kernel(const float* __restrict__ gt_ws, float* __restrict__ gcost, int D4, int H4, int W4, int H,
                                                                                 int W, float sh, float sw) {
  extern __shared__ float rowbuf[];  // [W] fine columns of this (plane, coarse row), rows already folded
  const int hc = blockIdx.x, d4 = blockIdx.y, b = blockIdx.z;
  const float* src = gt_ws + (((size_t)b * D4 + d4) * H) * W;
  // fine rows whose bilinear footprint includes coarse row hc: h0 == hc (weight lh0) and / or h1 == hc (weight lh1)
  // candidates: sh * h in (hc - 1, hc + 1), one row of slack on both sides (the exact membership test is repeated per row below)
  const int hlo = max(0, (int)floorf((hc - 1) / sh) - 1), hhi = min(H - 1, (int)ceilf((hc + 1) / sh) + 1);
  for (int w = threadIdx.x; w < W; w += kRegThreads) {
    float acc = 0.f;
    for (int h = hlo; h <= hhi; ++h) {
      const float hs = sh * h;
      const int h0 = (int)hs;
      const int h1 = h0 + (h0 < H4 - 1);
      const float lh1 = hs - h0, lh0 = 1.f - lh1;
      const float wgt = (h0 == hc ? lh0 : 0.f) + (h1 == hc ? lh1 : 0.f);
      if (wgt != 0.f) acc = fmaf(wgt, __ldg(src + (size_t)h * W + w), acc);
    }
    rowbuf[w] = acc;
  }
  __syncthreads();
  for (int wc = threadIdx.x; wc < W4; wc += kRegThreads) {
    const int wlo = max(0, (int)floorf((wc - 1) / sw) - 1), whi = min(W - 1, (int)ceilf((wc + 1) / sw) + 1);
    float acc = 0.f;
    for (int w = wlo; w <= whi; ++w) {
      const float ws = sw * w;
      const int w0 = (int)ws;
      const int w1 = w0 + (w0 < W4 - 1);
      const float lw1 = ws - w0, lw0 = 1.f - lw1;
      const float wgt = (w0 == wc ? lw0 : 0.f) + (w1 == wc ? lw1 : 0.f);
      acc = fmaf(wgt, rowbuf[w], acc);
    }
    gcost[(((size_t)b * D4 + d4) * H4 + hc) * W4 + wc] = acc;
  }
}

extern "C" size_t mode_disp_regress_backward_workspace_bytes(int B, int D4, int H4, int W4, int D, int H, int W) {
  const bool fast = D4 == 48 && D == 192 && W % kRegThreads == 0 && W4 * 4 == W && H4 * 4 == H && W4 >= 2 && H4 >= 2;
  return fast ? (size_t)B * D4 * H * W * sizeof(float) : 0;
}

extern "C" int mode_disp_regress_backward_ws(const float* cost, const float* grad_pred, float* grad_cost, void* workspace, int B, int D4, int H4, int W4, int D, int H, int W,
                                             void* stream) {
  MODE_CHECK_ARG(cost && grad_pred && grad_cost, "disp_regress_backward_ws: null pointer");
  if (workspace == nullptr || mode_disp_regress_backward_workspace_bytes(B, D4, H4, W4, D, H, W) == 0)
    return mode_disp_regress_backward(cost, grad_pred, grad_cost, B, D4, H4, W4, D, H, W, stream);
  MODE_CHECK_ARG(B > 0 && B < 65536, "disp_regress_backward_ws: bad batch");
  const float sh = (float)(H4 - 1) / (float)(H - 1), sw = (float)(W4 - 1) / (float)(W - 1);
  cudaStream_t s = (cudaStream_t)stream;
  dim3 g1((unsigned)((long long)H * (W / kRegThreads)), B);
  disp_regress_bwd192_pix_kernel<<<g1, kRegThreads, 0, s>>>(cost, grad_pred, (float*)workspace, H4, W4, H, W, sh, sw);
  MODE_CHECK_LAUNCH("disp_regress_backward_ws (per-pixel)");
  dim3 g2(H4, D4, B);
  disp_regress_bwd192_gather_kernel<<<g2, kRegThreads, (size_t)W * sizeof(float), s>>>((const float*)workspace, grad_cost, D4, H4, W4, H, W, sh, sw);
  MODE_CHECK_LAUNCH("disp_regress_backward_ws (gather)");
  return MODE_OK;
}

