// sphere_conv_f32.cu -- fp32 spherical convolution forward (parity mode, CUDA cores).
//
// Reference: sphere_conv_forward_cuda (sphere_conv_cuda.cpp:129-210) = per batch element
// sphere_im2col_gpu_kernel (sphere_conv_cuda_kernel.cu:195-262, bilinear :83-113) into a zero-filled
// (C*9, H*W) column buffer followed by addmm_ (cpp:191-196).  Here the gather and the contraction are one
// kernel: no column buffer (151 MB per element at Deep360 size) is ever written.
//   out[b,o,h,w] = sum_{c,k} W[o,c,k] * bilinear(x[b,c], pos[2k,h,w], pos[2k+1,h,w])
// with the reference's edge rules: the tap contributes only if h>-1 && w>-1 && h<H && w<W (kernel.cu:246) and
// each of the four corners is dropped individually when it lies outside the image (kernel.cu:97-107);
// there is no longitude wrap.  The sample is evaluated in the reference's expression order
// (w1*v1 + w2*v2 + w3*v3 + w4*v4, kernel.cu:109-111).
//
// Mapping: block = 32 pixels x NG output-channel groups of 32; each thread owns 32 accumulators.  Weights are
// staged in shared memory per 8-input-channel chunk as ws[ci][k][co] so a warp reads them as broadcast LDS.128.
#include "common.cuh"
using namespace mode;

constexpr int kCiChunk = 8;
constexpr int kCoPerThread = 32;

__global__ void __launch_bounds__(128) sphere_conv_f32_kernel(const float* __restrict__ x, const float* __restrict__ pos,
                                                              const float* __restrict__ wgt, const float* __restrict__ scale,
                                                              const float* __restrict__ shift, const float* __restrict__ residual,
                                                              float* __restrict__ out, int C, int H, int W, int Co, int KK, int relu) {
  extern __shared__ float ws[];  // [kCiChunk][KK][CoBlk]
  const int HW = H * W;
  const int CoBlk = blockDim.y * kCoPerThread;
  const int co_blk0 = blockIdx.y * CoBlk;
  const int b = blockIdx.z;
  const int pix = blockIdx.x * 32 + threadIdx.x;
  const bool active = pix < HW;
  const int p = active ? pix : HW - 1;
  const int co0 = threadIdx.y * kCoPerThread;
  const float* xb = x + (size_t)b * C * HW;
  const int tid = threadIdx.y * 32 + threadIdx.x, nthr = blockDim.y * 32;

  float acc[kCoPerThread];
#pragma unroll
  for (int i = 0; i < kCoPerThread; ++i) acc[i] = 0.f;

  for (int c0 = 0; c0 < C; c0 += kCiChunk) {
    __syncthreads();
    for (int e = tid; e < kCiChunk * KK * CoBlk; e += nthr) {
      int co = e % CoBlk, r = e / CoBlk;
      int k = r % KK, ci = r / KK;
      int cg = co_blk0 + co, c = c0 + ci;
      ws[e] = (cg < Co && c < C) ? __ldg(wgt + ((size_t)cg * C + c) * KK + k) : 0.f;
    }
    __syncthreads();
    for (int k = 0; k < KK; ++k) {
      const float h_im = __ldg(pos + (size_t)(2 * k) * HW + p);
      const float w_im = __ldg(pos + (size_t)(2 * k + 1) * HW + p);
      if (!(h_im > -1 && w_im > -1 && h_im < H && w_im < W)) continue;  // kernel.cu:246 (val = 0)
      const int h_low = (int)floorf(h_im), w_low = (int)floorf(w_im);
      const int h_high = h_low + 1, w_high = w_low + 1;
      const float lh = h_im - h_low, lw = w_im - w_low;
      const float hh = 1 - lh, hw = 1 - lw;
      const bool ok1 = h_low >= 0 && w_low >= 0, ok2 = h_low >= 0 && w_high <= W - 1;
      const bool ok3 = h_high <= H - 1 && w_low >= 0, ok4 = h_high <= H - 1 && w_high <= W - 1;
      const int o1 = h_low * W + w_low, o2 = h_low * W + w_high, o3 = h_high * W + w_low, o4 = h_high * W + w_high;
      const float w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
      const int cmax = min(kCiChunk, C - c0);
      for (int ci = 0; ci < cmax; ++ci) {
        const float* xc = xb + (size_t)(c0 + ci) * HW;
        const float v1 = ok1 ? __ldg(xc + o1) : 0.f;
        const float v2 = ok2 ? __ldg(xc + o2) : 0.f;
        const float v3 = ok3 ? __ldg(xc + o3) : 0.f;
        const float v4 = ok4 ? __ldg(xc + o4) : 0.f;
        const float val = (w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4);
        const float4* wr = reinterpret_cast<const float4*>(ws + ((size_t)ci * KK + k) * CoBlk + co0);
#pragma unroll
        for (int q = 0; q < kCoPerThread / 4; ++q) {
          const float4 wv = wr[q];
          acc[4 * q + 0] = fmaf(val, wv.x, acc[4 * q + 0]);
          acc[4 * q + 1] = fmaf(val, wv.y, acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(val, wv.z, acc[4 * q + 2]);
          acc[4 * q + 3] = fmaf(val, wv.w, acc[4 * q + 3]);
        }
      }
    }
  }
  if (!active) return;
#pragma unroll
  for (int i = 0; i < kCoPerThread; ++i) {
    const int co = co_blk0 + co0 + i;
    if (co < Co) {
      float y = acc[i];
      if (scale) y *= __ldg(scale + co);
      if (shift) y += __ldg(shift + co);
      const size_t o = ((size_t)b * Co + co) * HW + pix;
      if (residual) y += __ldg(residual + o);
      if (relu) y = fmaxf(y, 0.f);
      out[o] = y;
    }
  }
}

// ---- 3x3 kernels with C % 8 == 0 (every spherical layer of the model): register-tiled SGEMM with the gather fused into the A-tile
// staging.  The kernel above re-gathers every sample once per 32 output channels and feeds 32 FMAs from 8 broadcast LDS.128 -- it
// ran at 11.8 TFLOP/s (1.64 ms per 128->128 layer call of one pair), a quarter of the training step.  Here a block owns 128
// consecutive pixels x 128 output channels; per chunk of 8 input channels x 9 taps the 256 threads gather and blend the 72 x 128
// column tile ONCE into shared memory (reference expression order w1*v1+w2*v2+w3*v3+w4*v4, reference edge rules), stage the
// matching 72 x 128 weight tile, and every thread accumulates an 8 x 8 micro-tile: 4 LDS.128 per 64 FMAs.
constexpr int kTP = 128, kTC = 128, kCc = 8, kKK = 9, kKc = kCc * kKK;  // pixel tile, output-channel tile, channels per chunk, taps, K per chunk
constexpr int kTCp = kTC + 4;  // weight-tile row pitch: staging writes walk k (pitch 132 floats = 4 banks apart: 4-way instead of 32-way conflicts), rows stay 16-byte aligned

__global__ void __launch_bounds__(256, 2) sphere_conv_f32_tiled_kernel(const float* __restrict__ x, const float* __restrict__ pos, const float* __restrict__ wgt,
                                                                       const float* __restrict__ scale, const float* __restrict__ shift,
                                                                       const float* __restrict__ residual, float* __restrict__ out, int C, int H, int W, int Co, int relu) {
  extern __shared__ __align__(16) float sm_f[];
  float* a_s = sm_f;               // [kKc][kTP]  blended samples, k index = ci * 9 + tap
  float* b_s = sm_f + kKc * kTP;   // [kKc][kTCp] weights W[co][c0 + ci][tap]
  const int HW = H * W;
  const int tid = threadIdx.x;
  const int b = blockIdx.z, co_blk0 = blockIdx.y * kTC, pix0 = blockIdx.x * kTP;
  const float* xb = x + (size_t)b * C * HW;
  // gather role: pixel gp of the tile, channel half gh (4 of the chunk's 8 channels)
  const int gp = tid & (kTP - 1), gh = tid >> 7;
  const int gpix = min(pix0 + gp, HW - 1);
  // GEMM role: 8 pixels (tx*4.., 64 + tx*4..) x 8 output channels (ty*4.., 64 + ty*4..)
  const int tx = tid & 15, ty = tid >> 4;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  for (int c0 = 0; c0 < C; c0 += kCc) {
    __syncthreads();  // the previous chunk's tiles are consumed
    // ---- weights: b_s[ci*9 + k][co] = W[co_blk0 + co][c0 + ci][k]; a thread reads 72 consecutive floats of one output channel's row
    for (int e = tid; e < kKc * kTC; e += 256) {
      const int kk = e % kKc, co = e / kKc;
      const int cg = co_blk0 + co;
      b_s[kk * kTCp + co] = cg < Co ? __ldg(wgt + ((size_t)cg * C + c0) * kKK + kk) : 0.f;
    }
    // ---- samples: a_s[(ci)*9 + k][gp] for ci in this thread's channel half
#pragma unroll 1
    for (int k = 0; k < kKK; ++k) {
      const float h_im = __ldg(pos + (size_t)(2 * k) * HW + gpix);
      const float w_im = __ldg(pos + (size_t)(2 * k + 1) * HW + gpix);
      const bool tap = h_im > -1 && w_im > -1 && h_im < H && w_im < W;  // kernel.cu:246 (val = 0 otherwise)
      const int h_low = (int)floorf(h_im), w_low = (int)floorf(w_im);
      const int h_high = h_low + 1, w_high = w_low + 1;
      const float lh = h_im - h_low, lw = w_im - w_low;
      const float hh = 1 - lh, hw = 1 - lw;
      const bool ok1 = tap && h_low >= 0 && w_low >= 0, ok2 = tap && h_low >= 0 && w_high <= W - 1;
      const bool ok3 = tap && h_high <= H - 1 && w_low >= 0, ok4 = tap && h_high <= H - 1 && w_high <= W - 1;
      const int o1 = h_low * W + w_low, o2 = o1 + 1, o3 = o1 + W, o4 = o3 + 1;
      const float w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
#pragma unroll
      for (int cj = 0; cj < kCc / 2; ++cj) {
        const int ci = gh * (kCc / 2) + cj;
        const float* xc = xb + (size_t)(c0 + ci) * HW;
        const float v1 = ok1 ? __ldg(xc + o1) : 0.f;
        const float v2 = ok2 ? __ldg(xc + o2) : 0.f;
        const float v3 = ok3 ? __ldg(xc + o3) : 0.f;
        const float v4 = ok4 ? __ldg(xc + o4) : 0.f;
        a_s[(ci * kKK + k) * kTP + gp] = tap ? (w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < kKc; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(a_s + kk * kTP + tx * 4), a1 = *reinterpret_cast<const float4*>(a_s + kk * kTP + 64 + tx * 4);
      const float4 b0 = *reinterpret_cast<const float4*>(b_s + kk * kTCp + ty * 4), b1 = *reinterpret_cast<const float4*>(b_s + kk * kTCp + 64 + ty * 4);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w}, bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
  }
  // ---- epilogue: y = acc * scale + shift (+ residual) (ReLU); float4 along the pixel axis when the tile is full and aligned
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int co = co_blk0 + (j < 4 ? ty * 4 + j : 64 + ty * 4 + (j - 4));
    if (co >= Co) continue;
    const float sc = scale ? __ldg(scale + co) : 1.f, sh = shift ? __ldg(shift + co) : 0.f;
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      const int pl = pix0 + g * 64 + tx * 4;
      const size_t o = ((size_t)b * Co + co) * HW + pl;
      float y[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) y[i] = scale ? acc[g * 4 + i][j] * sc : acc[g * 4 + i][j];
#pragma unroll
      for (int i = 0; i < 4; ++i) y[i] = shift ? y[i] + sh : y[i];
      if (pl + 3 < HW && (HW & 3) == 0) {
        if (residual) {
          const float4 r = __ldg(reinterpret_cast<const float4*>(residual + o));
          y[0] += r.x, y[1] += r.y, y[2] += r.z, y[3] += r.w;
        }
        if (relu) y[0] = fmaxf(y[0], 0.f), y[1] = fmaxf(y[1], 0.f), y[2] = fmaxf(y[2], 0.f), y[3] = fmaxf(y[3], 0.f);
        *reinterpret_cast<float4*>(out + o) = make_float4(y[0], y[1], y[2], y[3]);
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (pl + i < HW) {
            float v = y[i];
            if (residual) v += __ldg(residual + o + i);
            out[o + i] = relu ? fmaxf(v, 0.f) : v;
          }
        }
      }
    }
  }
}

extern "C" int mode_sphere_conv_f32(const float* x, const float* pos, const float* w, const float* scale, const float* shift,
                                    const float* residual, float* out, int B, int C, int H, int W, int Co, int Kh, int Kw, int relu,
                                    void* stream) {
  MODE_CHECK_ARG(x && pos && w && out, "sphere_conv_f32: null pointer");
  MODE_CHECK_ARG(B > 0 && C > 0 && H > 0 && W > 0 && Co > 0 && Kh > 0 && Kw > 0, "sphere_conv_f32: bad shape");
  const int KK = Kh * Kw;
  if (KK == kKK && C % kCc == 0) {
    const size_t tsmem = (size_t)kKc * (kTP + kTCp) * sizeof(float);
    static thread_local bool tattr_dev[kMaxDevices] = {};
    bool& tattr = tattr_dev[current_device()];
    if (!tattr) {
      MODE_CHECK_CUDA(cudaFuncSetAttribute(sphere_conv_f32_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem), "sphere_conv_f32");
      tattr = true;
    }
    dim3 tgrid(ceil_div((long long)H * W, kTP), ceil_div(Co, kTC), B);
    sphere_conv_f32_tiled_kernel<<<tgrid, 256, tsmem, (cudaStream_t)stream>>>(x, pos, w, scale, shift, residual, out, C, H, W, Co, relu);
    MODE_CHECK_LAUNCH("sphere_conv_f32");
    return MODE_OK;
  }
  const int groups = ceil_div(Co, kCoPerThread);
  const int by = std::min(groups, 4);
  const int CoBlk = by * kCoPerThread;
  const size_t smem = (size_t)kCiChunk * KK * CoBlk * sizeof(float);
  MODE_CHECK_ARG(smem <= 200 * 1024, "sphere_conv_f32: kernel %dx%d too large", Kh, Kw);
  static thread_local size_t attr_dev[kMaxDevices] = {};  // the attribute is per device and per function
  size_t& attr = attr_dev[current_device()];
  if (smem > 48 * 1024 && smem > attr) {
    MODE_CHECK_CUDA(cudaFuncSetAttribute(sphere_conv_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "sphere_conv_f32");
    attr = smem;
  }
  dim3 grid(ceil_div((long long)H * W, 32), ceil_div(groups, by), B), block(32, by);
  sphere_conv_f32_kernel<<<grid, block, smem, (cudaStream_t)stream>>>(x, pos, w, scale, shift, residual, out, C, H, W, Co, KK, relu);
  MODE_CHECK_LAUNCH("sphere_conv_f32");
  return MODE_OK;
}
