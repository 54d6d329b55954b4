// costvol_conv.cu -- the cost volume and the first 3-D convolution on it (dres0[0]: 64 -> 32, 3x3x3, pad 1, + BN + ReLU) WITHOUT
// materialising the volume.
//
// Reference: cost volume mode_disparity.py:104-113, cost0 = dres0(cost) :115 with dres0[0] = convbn_3d(64, 32, 3, 1, 1) + ReLU
// (:66-67).
//
// The concatenation volume is a view of two 2-D feature maps:  cost[:, c, d, h, w] = ref[c, h, w] * [w >= d],
// cost[:, 32+c, d, h, w] = tgt[c, h, w-d] * [w >= d].  A 3x3x3 convolution over it therefore only ever multiplies weights with
// columns of ref / tgt.  With the per-(kd, kw) column responses (one GEMM per half, K = 3 kh x 32 c = 96, N = 9 x 32 = 288,
// computed by the caller with a library GEMM into fp32)
//     UR[kd,kw][o, h, x] = sum_{c,kh} W[o, c,    kd, kh, kw] * ref[c, h-1+kh, x]
//     UT[kd,kw][o, h, u] = sum_{c,kh} W[o, 32+c, kd, kh, kw] * tgt[c, h-1+kh, u]
// the convolution is EXACTLY
//     out[o, d, h, w] = sum over (kd, kw) with d' = d-1+kd in [0, D4), w' = w-1+kw in [0, W), w' >= d' of
//                       UR[kd,kw][o, h, w'] + UT[kd,kw][o, h, w'-d']
// and in the interior (1 <= d <= D4-2, d+2 <= w <= W-2), where all nine terms are present, it collapses to
//     out[o, d, h, w] = A0[o, h, w] + B0[o, h, w-d],   A0 = sum_{kd,kw} UR[kd,kw][.., w-1+kw],  B0 = sum UT[kd,kw][.., u+kw-kd]
// (two 2-D maps); for w <= d-3 every term is masked and out = 0.  So the 174 GFLOP / pair convolution over a 201 MB / pair
// volume becomes two small GEMMs and one HBM-write-bound kernel that reduces the A0 / B0 rows into shared memory and emits the
// 32-channel volume (2 shared-memory reads per output; in the diagonal band and on the volume's faces the few masked terms are
// subtracted again).  The same
// 16-bit-rounded operands are multiplied and everything is accumulated in fp32, so the result differs from the implicit-GEMM
// path only by fp32 summation order.
#include "common.cuh"
using namespace mode;

namespace {

// One CTA = one (b, h) row of the volume: the A0 / B0 rows (W x 32 fp32 each) are reduced from UR / UT into shared memory once,
// then the D4 x W x 32 outputs of the row are produced from there (each A0 / B0 element is reused ~D4 times; reading them
// from L2 per output made the kernel L2-read bound at 4x the bytes it writes).  Band / face voxels read UR / UT directly.
template <int FMT>
__global__ void __launch_bounds__(256, 3) costvol_conv_kernel(const float* __restrict__ ur, const float* __restrict__ ut, const float* __restrict__ scale,
                                                           const float* __restrict__ shift, uint16_t* __restrict__ out, int D4, int H, int W, int relu) {
  extern __shared__ __align__(16) float sm[];
  float* a0 = sm;                       // [W][32]
  float* b0 = sm + (size_t)W * 32;      // [W][32]
  float* a2 = sm + (size_t)W * 64;      // [W][32] A0 without its kd = 2 terms (far depth face)
  float* b2 = sm + (size_t)W * 96;      // [W][32] B0 without its kd = 2 terms
  const int b = blockIdx.x / H, h = blockIdx.x - b * H;
  const size_t prow = ((size_t)b * H + h) * W;  // pixel index of (b, h, 0)
  // A0[w][o] = sum_{kd,kw} UR[w-1+kw][kd*3+kw][o] (w-1+kw inside the row), B0[u][o] = sum UT[u+kw-kd][kd*3+kw][o]
  for (int e = threadIdx.x; e < W * 8; e += 256) {
    const int w = e >> 3, c4 = (e & 7) * 4;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), bb = a;
    // the kd = 0 terms are masked on the near face (d = 0: d' = -1), the kd = 2 terms on the far face (d = D4-1: d' = D4): keep their
    // sums, so that the two face planes need no second (DRAM-latency) pass over UR / UT -- ncu r02b: the kernel sat on long-scoreboard
    // stalls (17 per issue) of exactly those serialised corrections, its L2 hit rate 4 %
    float4 fa0 = a, fb0 = a, fa2 = a, fb2 = a;
#pragma unroll
    for (int kd = 0; kd < 3; ++kd)
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int wa = w - 1 + kw, wb = w + kw - kd;
        if (wa >= 0 && wa < W) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(ur + ((prow + wa) * 9 + kd * 3 + kw) * 32 + c4));
          a.x += v.x, a.y += v.y, a.z += v.z, a.w += v.w;
          if (kd == 0) fa0.x += v.x, fa0.y += v.y, fa0.z += v.z, fa0.w += v.w;
          if (kd == 2) fa2.x += v.x, fa2.y += v.y, fa2.z += v.z, fa2.w += v.w;
        }
        if (wb >= 0 && wb < W) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(ut + ((prow + wb) * 9 + kd * 3 + kw) * 32 + c4));
          bb.x += v.x, bb.y += v.y, bb.z += v.z, bb.w += v.w;
          if (kd == 0) fb0.x += v.x, fb0.y += v.y, fb0.z += v.z, fb0.w += v.w;
          if (kd == 2) fb2.x += v.x, fb2.y += v.y, fb2.z += v.z, fb2.w += v.w;
        }
      }
    // far face (d = D4-1), consumed below from shared memory: A-side and B-side sums without their kd = 2 terms
    *reinterpret_cast<float4*>(a2 + w * 32 + (c4 ^ ((w & 1) << 2))) = make_float4(a.x - fa2.x, a.y - fa2.y, a.z - fa2.z, a.w - fa2.w);
    *reinterpret_cast<float4*>(b2 + w * 32 + (c4 ^ ((w & 1) << 2))) = make_float4(bb.x - fb2.x, bb.y - fb2.y, bb.z - fb2.z, bb.w - fb2.w);
    // near face (d = 0, e = w: both sums belong to this thread): out = A0[w] + B0[w] - (kd = 0 terms); for 2 <= w <= W-2 nothing else
    // is masked there (d' = kd - 1 <= 1 <= w' and every column is inside the row)
    if (D4 >= 2 && w >= 2 && w <= W - 2) {
      float y[4] = {(a.x + bb.x) - (fa0.x + fb0.x), (a.y + bb.y) - (fa0.y + fb0.y), (a.z + bb.z) - (fa0.z + fb0.z), (a.w + bb.w) - (fa0.w + fb0.w)};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        y[i] = fmaf(y[i], scale ? __ldg(scale + c4 + i) : 1.f, shift ? __ldg(shift + c4 + i) : 0.f);
        y[i] = relu ? fmaxf(y[i], 0.f) : y[i];
      }
      *reinterpret_cast<uint2*>(out + ((((size_t)b * D4) * H + h) * W + w) * 32 + c4) = make_uint2(pack2<FMT>(y[0], y[1]), pack2<FMT>(y[2], y[3]));
    }
    // float4 slot q of row w is stored at slot q ^ (w & 1): the consumer below reads slots 2j and 2j + 1 (j = lane & 3) of two
    // CONSECUTIVE rows per quarter-warp -- un-swizzled those rows fall on the same banks (row pitch 128 B) and every LDS.128 was a 2-way
    // conflict (ncu r02: 48 % of this kernel's shared-memory wavefronts)
    *reinterpret_cast<float4*>(a0 + w * 32 + (c4 ^ ((w & 1) << 2))) = a;
    *reinterpret_cast<float4*>(b0 + w * 32 + (c4 ^ ((w & 1) << 2))) = bb;
  }
  __syncthreads();
  const int c8 = (threadIdx.x & 3) * 8;  // this thread's 8 output channels (fixed: the loop stride is a multiple of 4)
  float sc[8], sh[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) sc[i] = scale ? __ldg(scale + c8 + i) : 1.f, sh[i] = shift ? __ldg(shift + c8 + i) : 0.f;
  const int per_d = W * 4;
  // a thread keeps its pixel column w (and 8 channels) and walks the depth axis: A0[w] is depth-independent and stays in registers,
  // only the B0[w - d] row is read from shared memory per output
  for (int e = threadIdx.x; e < per_d; e += 256) {
    const int w = e >> 2;
    const float4* pa = reinterpret_cast<const float4*>(a0 + w * 32);
    const int qa = (c8 >> 2) ^ (w & 1);  // swizzled slot of channels c8..c8+3; the next four are slot ^ 1
    const float4 x0 = pa[qa], x1 = pa[qa ^ 1];
    const bool near_done = D4 >= 2 && w >= 2 && w <= W - 2;  // d = 0 was written by the reduction pass above
    for (int d = near_done ? 1 : 0; d < D4; ++d) {
      uint16_t* orow = out + ((((size_t)b * D4 + d) * H + h) * W) * 32;
      float acc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = 0.f;
      const int e_ = w - d;
      const bool far_fast = d == D4 - 1 && d >= 2 && e_ >= 2 && w <= W - 2;
      if (far_fast) {
        // far face away from the diagonal and the row ends: only the kd = 2 terms (d' = D4) are masked
        const float4* pa2 = reinterpret_cast<const float4*>(a2 + w * 32);
        const float4* pb2 = reinterpret_cast<const float4*>(b2 + e_ * 32);
        const int qb = (c8 >> 2) ^ (e_ & 1);
        const float4 u0 = pa2[qa], u1 = pa2[qa ^ 1], y0 = pb2[qb], y1 = pb2[qb ^ 1];
        acc[0] = u0.x + y0.x, acc[1] = u0.y + y0.y, acc[2] = u0.z + y0.z, acc[3] = u0.w + y0.w;
        acc[4] = u1.x + y1.x, acc[5] = u1.y + y1.y, acc[6] = u1.z + y1.z, acc[7] = u1.w + y1.w;
      }
      const bool mid = d >= 1 && d <= D4 - 2 && w >= 1;  // no depth-face term, w-1 inside the row
      auto sub8 = [&](const float* base, size_t pix, int kdkw) {
        const float4* q = reinterpret_cast<const float4*>(base + (pix * 9 + kdkw) * 32 + c8);
        const float4 z0 = __ldg(q), z1 = __ldg(q + 1);
        acc[0] -= z0.x, acc[1] -= z0.y, acc[2] -= z0.z, acc[3] -= z0.w, acc[4] -= z1.x, acc[5] -= z1.y, acc[6] -= z1.z, acc[7] -= z1.w;
      };
      auto add8 = [&](const float* base, size_t pix, int kdkw) {
        const float4* q = reinterpret_cast<const float4*>(base + (pix * 9 + kdkw) * 32 + c8);
        const float4 z0 = __ldg(q), z1 = __ldg(q + 1);
        acc[0] += z0.x, acc[1] += z0.y, acc[2] += z0.z, acc[3] += z0.w, acc[4] += z1.x, acc[5] += z1.y, acc[6] += z1.z, acc[7] += z1.w;
      };
      if (far_fast) {
        // done above
      } else if (e_ >= 0) {
        // all nine (kd, kw) terms, as far as A0 / B0 hold them ...
        const float4* pb = reinterpret_cast<const float4*>(b0 + e_ * 32);
        const int qb = (c8 >> 2) ^ (e_ & 1);
        const float4 y0 = pb[qb], y1 = pb[qb ^ 1];
        acc[0] = x0.x + y0.x, acc[1] = x0.y + y0.y, acc[2] = x0.z + y0.z, acc[3] = x0.w + y0.w;
        acc[4] = x1.x + y1.x, acc[5] = x1.y + y1.y, acc[6] = x1.z + y1.z, acc[7] = x1.w + y1.w;
        // ... minus the ones that are masked at this voxel.  Common cases first (pair index = kd*3 + kw):
        if (mid && e_ >= 2 && w <= W - 2) {
          // interior: nothing to remove
        } else if (mid && e_ >= 2 && d >= 2) {
          // w = W-1: the kw = 2 column lies outside the row; A0 skipped it, B0[e] (u + 2 - kd < W) did not
          sub8(ut, prow + e_ + 2, 2), sub8(ut, prow + e_ + 1, 5), sub8(ut, prow + e_, 8);
        } else if (mid && e_ == 1 && w <= W - 2) {
          sub8(ur, prow + w - 1, 6);  // (kd 2, kw 0): w' = w-1 < d' = d+1
        } else if (mid && e_ == 0 && w <= W - 2) {
          sub8(ur, prow + w - 1, 3), sub8(ur, prow + w - 1, 6), sub8(ur, prow + w, 7);  // (1,0), (2,0), (2,1)
        } else {
          // faces d = 0 / D4-1, w = 0, corners: generic
#pragma unroll
          for (int kd = 0; kd < 3; ++kd)
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
              const int dp = d - 1 + kd, wp = w - 1 + kw, up = e_ + kw - kd;
              const bool in_a = wp >= 0 && wp < W, in_b = up >= 0 && up < W;
              const bool valid = dp >= 0 && dp < D4 && in_a && wp >= dp;
              if (in_a && !valid) sub8(ur, prow + wp, kd * 3 + kw);
              if (in_b && !valid) sub8(ut, prow + up, kd * 3 + kw);
            }
        }
      } else if (e_ >= -2) {
        // left of the diagonal: at most three terms survive the mask w' >= d'
        if (mid && w <= W - 3 && e_ == -1) {
          add8(ur, prow + w, 1), add8(ut, prow, 1);          // (0,1): w' = w,   u' = 0
          add8(ur, prow + w + 1, 2), add8(ut, prow + 1, 2);  // (0,2): w' = w+1, u' = 1
          add8(ur, prow + w + 1, 5), add8(ut, prow, 5);      // (1,2): w' = w+1, u' = 0
        } else if (mid && w <= W - 3 && e_ == -2) {
          add8(ur, prow + w + 1, 2), add8(ut, prow, 2);      // (0,2): w' = w+1 = d' = d-1
        } else {
#pragma unroll
          for (int kd = 0; kd < 3; ++kd)
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
              const int dp = d - 1 + kd, wp = w - 1 + kw;
              if (dp < 0 || dp >= D4 || wp < 0 || wp >= W || wp < dp) continue;
              add8(ur, prow + wp, kd * 3 + kw), add8(ut, prow + wp - dp, kd * 3 + kw);
            }
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float y = fmaf(acc[i], sc[i], sh[i]);
        acc[i] = relu ? fmaxf(y, 0.f) : y;
      }
      const uint4 o = make_uint4(pack2<FMT>(acc[0], acc[1]), pack2<FMT>(acc[2], acc[3]), pack2<FMT>(acc[4], acc[5]), pack2<FMT>(acc[6], acc[7]));
      st_na_v4(orow + (size_t)w * 32 + c8, o);
    }
  }
}

// GEMM operand of the column responses: cols[p = (b,h,w)][kh*32 + c] = f[b, h-1+kh, w, c] (zero rows outside the image)
__global__ void __launch_bounds__(256) costvol_cols_kernel(const uint16_t* __restrict__ f, uint16_t* __restrict__ cols, int H, int W, long long nchunks) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= nchunks) return;
  const int ch = (int)(t % 12);  // 16-byte chunk of the 96-channel row: kh = ch / 4, channels (ch % 4) * 8 ..
  const long long p = t / 12;
  const int h = (int)((p / W) % H);
  const int kh = ch >> 2, hs = h - 1 + kh;
  uint4 v = make_uint4(0, 0, 0, 0);
  if (hs >= 0 && hs < H) v = __ldg(reinterpret_cast<const uint4*>(f + (p + (long long)(kh - 1) * W) * 32 + (ch & 3) * 8));
  *reinterpret_cast<uint4*>(cols + p * 96 + ch * 8) = v;
}

}  // namespace

extern "C" int mode_costvol_cols(const mode_h16* f, mode_h16* cols, int B, int H, int W, void* stream) {
  MODE_CHECK_ARG(f && cols && B > 0 && H > 0 && W > 0, "costvol_cols: bad arguments");
  const long long nchunks = (long long)B * H * W * 12;
  costvol_cols_kernel<<<(unsigned)ceil_div(nchunks, 256), 256, 0, (cudaStream_t)stream>>>(f, cols, H, W, nchunks);
  MODE_CHECK_LAUNCH("costvol_cols");
  return MODE_OK;
}

extern "C" int mode_costvol_conv_fused(const float* ur, const float* ut, const float* scale, const float* shift, mode_h16* out, int B, int D4, int H, int W, int relu,
                                       int fmt, void* stream) {
  MODE_CHECK_ARG(fmt == kFmtBF16 || fmt == kFmtFP16, "costvol_conv_fused: fmt must be 0 (bf16) or 1 (fp16)");
  MODE_CHECK_ARG(ur && ut && out, "costvol_conv_fused: null pointer");
  MODE_CHECK_ARG(B > 0 && D4 > 0 && H > 0 && W > 0, "costvol_conv_fused: bad shape");
  const size_t smem = (size_t)4 * W * 32 * sizeof(float);
  MODE_CHECK_ARG(smem <= 200 * 1024, "costvol_conv_fused: feature map too wide (W = %d)", W);
  static thread_local size_t attr_dev[kMaxDevices] = {};  // the attribute is per device and per function
  size_t& attr = attr_dev[current_device()];
  if (smem > 48 * 1024 && smem > attr) {
    MODE_CHECK_CUDA(cudaFuncSetAttribute(costvol_conv_kernel<kFmtBF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "costvol_conv_fused");
    MODE_CHECK_CUDA(cudaFuncSetAttribute(costvol_conv_kernel<kFmtFP16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "costvol_conv_fused");
    attr = smem;
  }
  cudaStream_t s = (cudaStream_t)stream;
  if (fmt == kFmtBF16)
    costvol_conv_kernel<kFmtBF16><<<B * H, 256, smem, s>>>(ur, ut, scale, shift, out, D4, H, W, relu);
  else
    costvol_conv_kernel<kFmtFP16><<<B * H, 256, smem, s>>>(ur, ut, scale, shift, out, D4, H, W, relu);
  MODE_CHECK_LAUNCH("costvol_conv_fused");
  return MODE_OK;
}
