// conv3d_f32.cu -- fp32 3x3x3 conv / strided conv / transposed conv with fused BN-affine + residual + ReLU
// (parity mode, CUDA cores).  Reference: nn.Conv3d / nn.ConvTranspose3d + BatchNorm3d in
// models/submodule.py:20-22 and models/mode_disparity.py:15-25,66-80; dataflow :27-46,:115-129.
//
//   mode 0: out[o] = sum x[o - 1 + k]        (stride 1, pad 1)
//   mode 1: out[o] = sum x[2o - 1 + k]       (stride 2, pad 1)
//   mode 2: out[o] = sum x[(o + 1 - k)/2]    (transposed, stride 2, pad 1, output_padding 1; only even o+1-k)
//
// Mapping: block = 32 consecutive output w x NG groups of 8 output channels, one (b, od, oh) row per block.y.
// Weights staged in smem per 8-input-channel chunk as ws[ci][tap][co] (broadcast LDS.128 per warp).
#include "common.cuh"
using namespace mode;

constexpr int kCi3 = 8;
constexpr int kCoT = 8;

__global__ void __launch_bounds__(256) conv3d_f32_kernel(const float* __restrict__ x, const float* __restrict__ wgt,
                                                         const float* __restrict__ scale, const float* __restrict__ shift,
                                                         const float* __restrict__ residual, float* __restrict__ out, int Ci, int Co,
                                                         int Di, int Hi, int Wi, int Do, int Ho, int Wo, int mode, int relu) {
  extern __shared__ float ws[];  // [kCi3][27][CoBlk]
  const int CoBlk = blockDim.y * kCoT;
  const int co_blk0 = blockIdx.y * CoBlk;
  const int wtiles = (Wo + 31) >> 5;
  int row = blockIdx.x / wtiles;
  const int wt = blockIdx.x - row * wtiles;
  const int oh = row % Ho;
  row /= Ho;
  const int od = row % Do;
  const int b = row / Do;
  const int ow = wt * 32 + threadIdx.x;
  const bool active = ow < Wo;
  const int co0 = threadIdx.y * kCoT;
  const size_t DHWi = (size_t)Di * Hi * Wi;
  const float* xb = x + (size_t)b * Ci * DHWi;
  const int tid = threadIdx.y * 32 + threadIdx.x, nthr = blockDim.y * 32;

  float acc[kCoT];
#pragma unroll
  for (int i = 0; i < kCoT; ++i) acc[i] = 0.f;

  for (int c0 = 0; c0 < Ci; c0 += kCi3) {
    __syncthreads();
    for (int e = tid; e < kCi3 * 27 * CoBlk; e += nthr) {
      int co = e % CoBlk, r = e / CoBlk;
      int k = r % 27, ci = r / 27;
      int cg = co_blk0 + co, c = c0 + ci;
      float v = 0.f;
      if (cg < Co && c < Ci) v = (mode == 2) ? __ldg(wgt + ((size_t)c * Co + cg) * 27 + k) : __ldg(wgt + ((size_t)cg * Ci + c) * 27 + k);
      ws[e] = v;
    }
    __syncthreads();
    if (!active) continue;
    const int cmax = min(kCi3, Ci - c0);
    for (int kd = 0; kd < 3; ++kd) {
      int id;
      bool vd;
      if (mode == 2) {
        int n = od + 1 - kd;
        vd = n >= 0 && !(n & 1) && (n >> 1) < Di;
        id = n >> 1;
      } else {
        id = od * (mode + 1) - 1 + kd;
        vd = id >= 0 && id < Di;
      }
      if (!vd) continue;
      for (int kh = 0; kh < 3; ++kh) {
        int ih;
        bool vh;
        if (mode == 2) {
          int n = oh + 1 - kh;
          vh = n >= 0 && !(n & 1) && (n >> 1) < Hi;
          ih = n >> 1;
        } else {
          ih = oh * (mode + 1) - 1 + kh;
          vh = ih >= 0 && ih < Hi;
        }
        if (!vh) continue;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          int iw;
          bool vw;
          if (mode == 2) {
            int n = ow + 1 - kw;
            vw = n >= 0 && !(n & 1) && (n >> 1) < Wi;
            iw = n >> 1;
          } else {
            iw = ow * (mode + 1) - 1 + kw;
            vw = iw >= 0 && iw < Wi;
          }
          if (!vw) continue;
          const int tap = (kd * 3 + kh) * 3 + kw;
          const float* xp = xb + (size_t)c0 * DHWi + ((size_t)id * Hi + ih) * Wi + iw;
          for (int ci = 0; ci < cmax; ++ci) {
            const float v = __ldg(xp + (size_t)ci * DHWi);
            const float4* wr = reinterpret_cast<const float4*>(ws + ((size_t)ci * 27 + tap) * CoBlk + co0);
            const float4 wa = wr[0], wb = wr[1];
            acc[0] = fmaf(v, wa.x, acc[0]);
            acc[1] = fmaf(v, wa.y, acc[1]);
            acc[2] = fmaf(v, wa.z, acc[2]);
            acc[3] = fmaf(v, wa.w, acc[3]);
            acc[4] = fmaf(v, wb.x, acc[4]);
            acc[5] = fmaf(v, wb.y, acc[5]);
            acc[6] = fmaf(v, wb.z, acc[6]);
            acc[7] = fmaf(v, wb.w, acc[7]);
          }
        }
      }
    }
  }
  if (!active) return;
  const size_t DHWo = (size_t)Do * Ho * Wo;
#pragma unroll
  for (int i = 0; i < kCoT; ++i) {
    const int co = co_blk0 + co0 + i;
    if (co < Co) {
      float y = acc[i];
      if (scale) y *= __ldg(scale + co);
      if (shift) y += __ldg(shift + co);
      const size_t o = ((size_t)b * Co + co) * DHWo + ((size_t)od * Ho + oh) * Wo + ow;
      if (residual) y += __ldg(residual + o);
      if (relu) y = fmaxf(y, 0.f);
      out[o] = y;
    }
  }
}

extern "C" int mode_conv3d_f32(const float* x, const float* w, const float* scale, const float* shift, const float* residual, float* out,
                               int B, int Ci, int Co, int Di, int Hi, int Wi, int mode, int relu, void* stream) {
  MODE_CHECK_ARG(x && w && out, "conv3d_f32: null pointer");
  MODE_CHECK_ARG(B > 0 && Ci > 0 && Co > 0 && Di > 0 && Hi > 0 && Wi > 0, "conv3d_f32: bad shape");
  MODE_CHECK_ARG(mode >= 0 && mode <= 2, "conv3d_f32: mode %d not in {0,1,2}", mode);
  int Do, Ho, Wo;
  if (mode == 0) {
    Do = Di, Ho = Hi, Wo = Wi;
  } else if (mode == 1) {
    Do = (Di - 1) / 2 + 1, Ho = (Hi - 1) / 2 + 1, Wo = (Wi - 1) / 2 + 1;
  } else {
    Do = 2 * Di, Ho = 2 * Hi, Wo = 2 * Wi;
  }
  const int groups = ceil_div(Co, kCoT);
  const int by = std::min(groups, 8);
  const size_t smem = (size_t)kCi3 * 27 * by * kCoT * sizeof(float);
  static thread_local size_t attr_dev[kMaxDevices] = {};  // the attribute is per device and per function
  size_t& attr = attr_dev[current_device()];
  if (smem > 48 * 1024 && smem > attr) {
    MODE_CHECK_CUDA(cudaFuncSetAttribute(conv3d_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "conv3d_f32");
    attr = smem;
  }
  const long long nblk = (long long)B * Do * Ho * ceil_div(Wo, 32);
  MODE_CHECK_ARG(nblk < 2147483647LL, "conv3d_f32: grid too large");
  dim3 grid((unsigned)nblk, ceil_div(groups, by)), block(32, by);
  conv3d_f32_kernel<<<grid, block, smem, (cudaStream_t)stream>>>(x, w, scale, shift, residual, out, Ci, Co, Di, Hi, Wi, Do, Ho, Wo, mode, relu);
  MODE_CHECK_LAUNCH("conv3d_f32");
  return MODE_OK;
}
