// cost_volume.cu -- PSMNet-style concatenation cost volume (reference: models/mode_disparity.py:104-113).
//
// Pure data movement, HBM-write bound: per pair at 1024x512/D=192 the kernel reads 2 x 4.19 MB (L2 resident
// after first touch) and writes 402.65 MB (fp32 NCDHW) or 201.3 MB (bf16 NDHWC).  128-bit stores, fully coalesced in
// both layouts, four per thread in flight.  Shift indices are integers -> bit-exact.
#include "common.cuh"
using namespace mode;

// fp32 NCDHW: one CTA = one (b, c2, i) plane of H x W floats, a thread = groups of 4 consecutive w, 4 groups in flight.
__global__ void __launch_bounds__(256) cost_volume_f32_kernel(const float* __restrict__ ref, const float* __restrict__ tgt,
                                                              float* __restrict__ cost, int C, int H, int W, int D4, long long planes) {
  const int W4 = W >> 2, per_plane = H * W4;
  for (long long pl = blockIdx.x; pl < planes; pl += gridDim.x) {
    const int i = (int)(pl % D4);
    const long long r = pl / D4;
    const int c2 = (int)(r % (2 * C)), b = (int)(r / (2 * C));
    const bool left = c2 < C;
    const float* src = left ? ref + ((size_t)b * C + c2) * H * W : tgt + ((size_t)b * C + (c2 - C)) * H * W;
    float4* out = reinterpret_cast<float4*>(cost + (size_t)pl * H * W);
    for (int e0 = threadIdx.x; e0 < per_plane; e0 += 4 * 256) {
      float4 v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int e = e0 + k * 256;
        v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (e < per_plane) {
          const int h = e / W4, w = (e - h * W4) << 2;
          const float* s = src + (size_t)h * W + w;
          if (left) {
            const float4 q = *reinterpret_cast<const float4*>(s);
            v[k].x = (w + 0 >= i) ? q.x : 0.f, v[k].y = (w + 1 >= i) ? q.y : 0.f, v[k].z = (w + 2 >= i) ? q.z : 0.f, v[k].w = (w + 3 >= i) ? q.w : 0.f;
          } else {
            v[k].x = (w + 0 >= i) ? __ldg(s + 0 - i) : 0.f, v[k].y = (w + 1 >= i) ? __ldg(s + 1 - i) : 0.f;
            v[k].z = (w + 2 >= i) ? __ldg(s + 2 - i) : 0.f, v[k].w = (w + 3 >= i) ? __ldg(s + 3 - i) : 0.f;
          }
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int e = e0 + k * 256;
        if (e < per_plane) st_na_v4(out + e, *reinterpret_cast<uint4*>(&v[k]));
      }
    }
  }
}

// 16-bit NDHWC: voxel = 2C channels; one CTA = one (b, disparity i, h) row of W voxels, a thread = 16-byte chunks (8
// channels) of that row, 4 loads in flight before the first store (no per-element index division: the kernel is a pure
// HBM-write stream, 201 MB per pair, and every instruction not a load or a store only lowers the bytes in flight).
__global__ void __launch_bounds__(256) cost_volume_bf16_kernel(const uint16_t* __restrict__ ref, const uint16_t* __restrict__ tgt,
                                                               uint16_t* __restrict__ cost, int C, int H, int W, int D4, long long rows) {
  const int chunks = (2 * C) >> 3, half = C >> 3;  // 16-byte chunks per voxel / per feature vector
  const int per_row = W * chunks;
  for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
    const int h = (int)(row % H);
    const long long r = row / H;
    const int i = (int)(r % D4), b = (int)(r / D4);
    const uint4* rrow = reinterpret_cast<const uint4*>(ref + ((size_t)b * H + h) * W * C);
    const uint4* trow = reinterpret_cast<const uint4*>(tgt + ((size_t)b * H + h) * W * C);
    uint4* out = reinterpret_cast<uint4*>(cost + (size_t)row * W * 2 * C);
    for (int e0 = threadIdx.x; e0 < per_row; e0 += 4 * 256) {
      uint4 v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int e = e0 + k * 256;
        v[k] = make_uint4(0, 0, 0, 0);
        if (e < per_row) {
          const int w = e / chunks, ch = e - w * chunks;
          if (w >= i) v[k] = ch < half ? __ldg(rrow + w * half + ch) : __ldg(trow + (w - i) * half + (ch - half));
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int e = e0 + k * 256;
        if (e < per_row) st_na_v4(out + e, v[k]);
      }
    }
  }
}

extern "C" int mode_cost_volume_f32(const float* ref, const float* tgt, float* cost, int B, int C, int H, int W, int D4, void* stream) {
  MODE_CHECK_ARG(ref && tgt && cost, "cost_volume_f32: null pointer");
  MODE_CHECK_ARG(B > 0 && C > 0 && H > 0 && W > 0 && D4 > 0, "cost_volume_f32: bad shape B=%d C=%d H=%d W=%d D4=%d", B, C, H, W, D4);
  MODE_CHECK_ARG(W % 4 == 0, "cost_volume_f32: W (%d) must be a multiple of 4", W);
  const long long planes = (long long)B * 2 * C * D4;
  const int blocks = (int)std::min<long long>(planes, (long long)kNumSMs * 16);
  cost_volume_f32_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(ref, tgt, cost, C, H, W, D4, planes);
  MODE_CHECK_LAUNCH("cost_volume_f32");
  return MODE_OK;
}

extern "C" int mode_cost_volume_16(const mode_h16* ref, const mode_h16* tgt, mode_h16* cost, int B, int C, int H, int W, int D4,
                                     void* stream) {
  MODE_CHECK_ARG(ref && tgt && cost, "cost_volume_16: null pointer");
  MODE_CHECK_ARG(B > 0 && C > 0 && H > 0 && W > 0 && D4 > 0, "cost_volume_16: bad shape B=%d C=%d H=%d W=%d D4=%d", B, C, H, W, D4);
  MODE_CHECK_ARG(C % 8 == 0, "cost_volume_16: C (%d) must be a multiple of 8", C);
  const long long rows = (long long)B * D4 * H;
  const int blocks = (int)std::min<long long>(rows, (long long)kNumSMs * 16);
  cost_volume_bf16_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(ref, tgt, cost, C, H, W, D4, rows);
  MODE_CHECK_LAUNCH("cost_volume_16");
  return MODE_OK;
}

// ---- backward (training): the volume is a set of shifted views, so its gradient is a gather-sum over the D/4 shifts
//   grad_ref[b,c,h,w]  = sum_{i <= w, i < D4}       g[b, c,     i, h, w]
//   grad_tgt[b,c,h,w'] = sum_{i < D4, w' + i < W}   g[b, C + c, i, h, w' + i]
// (reference: autograd through the slice assignments of models/mode_disparity.py:104-113).  One thread per output element, the D/4
// reads of a warp are coalesced rows of the (B,2C,D4,H,W) gradient; deterministic (no atomics), fixed summation order i = 0..D4-1.
__global__ void __launch_bounds__(256) cost_volume_bwd_f32_kernel(const float* __restrict__ g, float* __restrict__ gref, float* __restrict__ gtgt, int C, int H, int W,
                                                                  int D4, long long n) {
  const size_t plane = (size_t)H * W;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(e % W);
    long long r = e / W;
    const int h = (int)(r % H);
    r /= H;
    const int c2 = (int)(r % (2 * C)), b = (int)(r / (2 * C));
    const float* gp = g + (((size_t)b * 2 * C + c2) * D4) * plane + (size_t)h * W + w;
    float acc = 0.f;
    if (c2 < C) {
      const int imax = min(w, D4 - 1);
      for (int i = 0; i <= imax; ++i) acc += __ldg(gp + (size_t)i * plane);
      gref[(((size_t)b * C + c2) * H + h) * W + w] = acc;
    } else {
      const int imax = min(D4 - 1, W - 1 - w);
      for (int i = 0; i <= imax; ++i) acc += __ldg(gp + (size_t)i * plane + i);
      gtgt[(((size_t)b * C + (c2 - C)) * H + h) * W + w] = acc;
    }
  }
}

extern "C" int mode_cost_volume_backward_f32(const float* grad_cost, float* grad_ref, float* grad_tgt, int B, int C, int H, int W, int D4, void* stream) {
  MODE_CHECK_ARG(grad_cost && grad_ref && grad_tgt, "cost_volume_backward_f32: null pointer");
  MODE_CHECK_ARG(B > 0 && C > 0 && H > 0 && W > 0 && D4 > 0, "cost_volume_backward_f32: bad shape");
  const long long n = (long long)B * 2 * C * H * W;
  const int blocks = (int)std::min<long long>((n + 255) / 256, (long long)kNumSMs * 32);
  cost_volume_bwd_f32_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(grad_cost, grad_ref, grad_tgt, C, H, W, D4, n);
  MODE_CHECK_LAUNCH("cost_volume_backward_f32");
  return MODE_OK;
}
