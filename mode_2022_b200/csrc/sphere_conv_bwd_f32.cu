// sphere_conv_bwd_f32.cu -- fp32 spherical convolution backward (training path, CUDA cores).
//
// Reference: sphere_conv_backward_cuda (sphere_conv_cuda.cpp:213-336): per batch element
//   columns = W^T . grad_output            (addmm_, cpp:276-279)   -> sphere_col2im_gpu_kernel (kernel.cu:293-356, atomicAdd
//                                                                     of get_gradient_weight (:128-152) x column)   = grad_input
//   columns = sphere_im2col(input)         (kernel.cu:195-262)     -> grad_weight += grad_output . columns^T (cpp:296-303)
//   grad_bias += grad_output . ones                                   (cpp:304-310)
// i.e. two 151 MB column buffers per element at Deep360 size.  Here each gradient is ONE kernel and no column buffer exists:
//   dgrad: cols[c] = sum_o W[o,c,k] * gout[b,o,pix] is formed in registers per (pixel, tap, 32 input channels) and scattered
//          through the transposed bilinear stencil -- the same four corners and weights as the forward gather, with the
//          forward's edge rules (tap dropped unless -1 < h < H and -1 < w < W, each corner dropped outside the image);
//          get_gradient_weight evaluates exactly those weights, and the reference's (int) truncation + |d| < 1 test selects
//          exactly those corners.
//   wgrad: gW[o,c,k] = sum_{b,pix} gout[b,o,pix] * bilinear(x[b,c], pos[k,pix]): 64 x 64 register-tiled GEMM over the pixel
//          axis whose B operand is gathered on the fly; partial tiles are combined with atomicAdd.
// Like the reference, both ACCUMULATE into caller-zeroed gradients (sphere_conv.py:62-64) and the scatter uses fp32 atomics
// (summation order, hence the last bits, is not deterministic -- SURVEY.md section 8c fixture rule 4).
#include "common.cuh"
using namespace mode;

namespace {

struct Stencil {
  int o1, o2, o3, o4;
  float w1, w2, w3, w4;  // 0 for dropped corners / dropped tap
};
__device__ __forceinline__ Stencil make_stencil(float h_im, float w_im, int H, int W) {
  Stencil s;
  s.o1 = s.o2 = s.o3 = s.o4 = 0;
  s.w1 = s.w2 = s.w3 = s.w4 = 0.f;
  if (!(h_im > -1 && w_im > -1 && h_im < H && w_im < W)) return s;  // kernel.cu:246 / :132-135
  const int h_low = (int)floorf(h_im), w_low = (int)floorf(w_im);
  const int h_high = h_low + 1, w_high = w_low + 1;
  const float lh = h_im - h_low, lw = w_im - w_low;
  const float hh = 1 - lh, hw = 1 - lw;
  if (h_low >= 0 && w_low >= 0) s.o1 = h_low * W + w_low, s.w1 = hh * hw;
  if (h_low >= 0 && w_high <= W - 1) s.o2 = h_low * W + w_high, s.w2 = hh * lw;
  if (h_high <= H - 1 && w_low >= 0) s.o3 = h_high * W + w_low, s.w3 = lh * hw;
  if (h_high <= H - 1 && w_high <= W - 1) s.o4 = h_high * W + w_high, s.w4 = lh * lw;
  return s;
}

constexpr int kCPerThread = 32;

// ---- dgrad: block = 32 pixels x (blockDim.y groups of 32 input channels); per tap the (Co x Cblk) weight slice sits in smem
__global__ void __launch_bounds__(128) sphere_dgrad_f32_kernel(const float* __restrict__ gout, const float* __restrict__ pos, const float* __restrict__ wgt,
                                                               float* __restrict__ gin, int C, int H, int W, int Co, int KK) {
  extern __shared__ float ws[];  // [Co][Cblk]
  const int HW = H * W;
  const int Cblk = blockDim.y * kCPerThread;
  const int c_blk0 = blockIdx.y * Cblk;
  const int b = blockIdx.z;
  const int pix = blockIdx.x * 32 + threadIdx.x;
  const bool active = pix < HW;
  const int p = active ? pix : HW - 1;
  const int c0 = threadIdx.y * kCPerThread;
  const int tid = threadIdx.y * 32 + threadIdx.x, nthr = blockDim.y * 32;
  const float* gb = gout + (size_t)b * Co * HW + p;
  float* gi = gin + (size_t)b * C * HW;
  for (int k = 0; k < KK; ++k) {
    __syncthreads();
    for (int e = tid; e < Co * Cblk; e += nthr) {
      const int c = e % Cblk, o = e / Cblk;
      ws[e] = (c_blk0 + c < C) ? __ldg(wgt + ((size_t)o * C + c_blk0 + c) * KK + k) : 0.f;
    }
    __syncthreads();
    const Stencil st = make_stencil(__ldg(pos + (size_t)(2 * k) * HW + p), __ldg(pos + (size_t)(2 * k + 1) * HW + p), H, W);
    if (!active || (st.w1 == 0.f && st.w2 == 0.f && st.w3 == 0.f && st.w4 == 0.f)) continue;  // block-uniform barriers are above
    float acc[kCPerThread];
#pragma unroll
    for (int i = 0; i < kCPerThread; ++i) acc[i] = 0.f;
    for (int o = 0; o < Co; ++o) {
      const float g = __ldg(gb + (size_t)o * HW);
      const float4* wr = reinterpret_cast<const float4*>(ws + (size_t)o * Cblk + c0);
#pragma unroll
      for (int q = 0; q < kCPerThread / 4; ++q) {
        const float4 wv = wr[q];
        acc[4 * q + 0] = fmaf(g, wv.x, acc[4 * q + 0]);
        acc[4 * q + 1] = fmaf(g, wv.y, acc[4 * q + 1]);
        acc[4 * q + 2] = fmaf(g, wv.z, acc[4 * q + 2]);
        acc[4 * q + 3] = fmaf(g, wv.w, acc[4 * q + 3]);
      }
    }
#pragma unroll
    for (int i = 0; i < kCPerThread; ++i) {
      const int c = c_blk0 + c0 + i;
      if (c < C) {
        float* gc = gi + (size_t)c * HW;
        if (st.w1 != 0.f) atomicAdd(gc + st.o1, st.w1 * acc[i]);
        if (st.w2 != 0.f) atomicAdd(gc + st.o2, st.w2 * acc[i]);
        if (st.w3 != 0.f) atomicAdd(gc + st.o3, st.w3 * acc[i]);
        if (st.w4 != 0.f) atomicAdd(gc + st.o4, st.w4 * acc[i]);
      }
    }
  }
}

// ---- wgrad: block = 64 (o) x 64 (c) tile of one tap over one pixel segment; 256 threads x 4x4 accumulators
constexpr int kWgTile = 64, kWgPix = 32, kWgPad = 68;
__global__ void __launch_bounds__(256) sphere_wgrad_f32_kernel(const float* __restrict__ x, const float* __restrict__ pos, const float* __restrict__ gout,
                                                               float* __restrict__ gw, int C, int H, int W, int Co, int KK, int seg_pix, int segs) {
  __shared__ __align__(16) float g_s[kWgPix][kWgPad];  // [pixel][o]
  __shared__ __align__(16) float v_s[kWgPix][kWgPad];  // [pixel][c]
  const int HW = H * W;
  const int b = blockIdx.x / segs, seg = blockIdx.x - b * segs;
  const int k = blockIdx.y;
  const int ctiles = (C + kWgTile - 1) / kWgTile;
  const int o_blk0 = (blockIdx.z / ctiles) * kWgTile, c_blk0 = (blockIdx.z % ctiles) * kWgTile;
  const int t = threadIdx.x, to = t >> 4, tc = t & 15;
  const float* xb = x + (size_t)b * C * HW;
  const float* gb = gout + (size_t)b * Co * HW;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int p_end = min((seg + 1) * seg_pix, HW);
  for (int p0 = seg * seg_pix; p0 < p_end; p0 += kWgPix) {
    __syncthreads();
    // gout tile: 64 o x 32 pixels (coalesced along pixels), stored [pixel][o]
    for (int e = t; e < kWgTile * kWgPix; e += 256) {
      const int px = e & 31, o = e >> 5;
      const int p = p0 + px;
      g_s[px][o] = (p < p_end && o_blk0 + o < Co) ? __ldg(gb + (size_t)(o_blk0 + o) * HW + p) : 0.f;
    }
    // sampled input tile: 64 c x 32 pixels (the forward's bilinear sample, same expression order), stored [pixel][c]
    {
      const int px = t & 31;
      const int p = p0 + px;
      Stencil st = make_stencil(0.f, 0.f, 0, 0);
      if (p < p_end) st = make_stencil(__ldg(pos + (size_t)(2 * k) * HW + p), __ldg(pos + (size_t)(2 * k + 1) * HW + p), H, W);
      for (int c = t >> 5; c < kWgTile; c += 8) {
        float val = 0.f;
        if (c_blk0 + c < C) {
          const float* xc = xb + (size_t)(c_blk0 + c) * HW;
          const float v1 = st.w1 != 0.f ? __ldg(xc + st.o1) : 0.f, v2 = st.w2 != 0.f ? __ldg(xc + st.o2) : 0.f;
          const float v3 = st.w3 != 0.f ? __ldg(xc + st.o3) : 0.f, v4 = st.w4 != 0.f ? __ldg(xc + st.o4) : 0.f;
          val = (st.w1 * v1 + st.w2 * v2 + st.w3 * v3 + st.w4 * v4);
        }
        v_s[px][c] = val;
      }
    }
    __syncthreads();
#pragma unroll 8
    for (int px = 0; px < kWgPix; ++px) {
      const float4 g = *reinterpret_cast<const float4*>(&g_s[px][4 * to]);
      const float4 v = *reinterpret_cast<const float4*>(&v_s[px][4 * tc]);
      const float ga[4] = {g.x, g.y, g.z, g.w}, va[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ga[i], va[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int o = o_blk0 + 4 * to + i;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = c_blk0 + 4 * tc + j;
      if (o < Co && c < C) atomicAdd(gw + ((size_t)o * C + c) * KK + k, acc[i][j]);
    }
  }
}

// ---- grad_bias[o] += sum_{b,pix} gout[b,o,pix]
__global__ void __launch_bounds__(256) sphere_bgrad_f32_kernel(const float* __restrict__ gout, float* __restrict__ gbias, int B, int Co, int HW) {
  const int o = blockIdx.x;
  float s = 0.f;
  for (int b = blockIdx.y; b < B; b += gridDim.y)
    for (int p = threadIdx.x; p < HW; p += 256) s += __ldg(gout + ((size_t)b * Co + o) * HW + p);
  __shared__ float red[8];
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int i = 0; i < 8; ++i) tot += red[i];
    atomicAdd(gbias + o, tot);
  }
}

}  // namespace

extern "C" int mode_sphere_conv_backward_f32(const float* x, const float* pos, const float* w, const float* grad_out, float* grad_in, float* grad_w,
                                             float* grad_bias, int B, int C, int H, int W, int Co, int Kh, int Kw, void* stream) {
  MODE_CHECK_ARG(pos && grad_out, "sphere_conv_backward_f32: null pointer");
  MODE_CHECK_ARG(B > 0 && C > 0 && H > 0 && W > 0 && Co > 0 && Kh > 0 && Kw > 0, "sphere_conv_backward_f32: bad shape");
  MODE_CHECK_ARG(!grad_in || w, "sphere_conv_backward_f32: grad_in needs the weights");
  MODE_CHECK_ARG(!grad_w || x, "sphere_conv_backward_f32: grad_w needs the input");
  const int KK = Kh * Kw, HW = H * W;
  cudaStream_t s = (cudaStream_t)stream;
  if (grad_in) {
    const int groups = ceil_div(C, kCPerThread);
    const int by = std::min(groups, 4);
    const size_t smem = (size_t)Co * by * kCPerThread * sizeof(float);
    MODE_CHECK_ARG(smem <= 200 * 1024, "sphere_conv_backward_f32: Co = %d too large", Co);
    static thread_local size_t attr_dev[kMaxDevices] = {};  // the attribute is per device and per function
  size_t& attr = attr_dev[current_device()];
    if (smem > 48 * 1024 && smem > attr) {
      MODE_CHECK_CUDA(cudaFuncSetAttribute(sphere_dgrad_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "sphere_conv_backward_f32");
      attr = smem;
    }
    dim3 grid(ceil_div((long long)HW, 32), ceil_div(groups, by), B), block(32, by);
    sphere_dgrad_f32_kernel<<<grid, block, smem, s>>>(grad_out, pos, w, grad_in, C, H, W, Co, KK);
    MODE_CHECK_LAUNCH("sphere_conv_backward_f32 (dgrad)");
  }
  if (grad_w) {
    // pixel segments: enough blocks to fill the GPU, few enough that the atomic combine stays small
    const int tiles = ceil_div(Co, kWgTile) * ceil_div(C, kWgTile);
    int segs = std::max(1, std::min(ceil_div(HW, 256), ceil_div(4 * kNumSMs, B * KK * tiles)));
    int seg_pix = ceil_div(ceil_div(HW, segs), kWgPix) * kWgPix;
    segs = ceil_div(HW, seg_pix);
    dim3 grid(B * segs, KK, tiles);
    sphere_wgrad_f32_kernel<<<grid, 256, 0, s>>>(x, pos, grad_out, grad_w, C, H, W, Co, KK, seg_pix, segs);
    MODE_CHECK_LAUNCH("sphere_conv_backward_f32 (wgrad)");
  }
  if (grad_bias) {
    dim3 grid(Co, std::min(B, 8));
    sphere_bgrad_f32_kernel<<<grid, 256, 0, s>>>(grad_out, grad_bias, B, Co, HW);
    MODE_CHECK_LAUNCH("sphere_conv_backward_f32 (bias)");
  }
  return MODE_OK;
}
