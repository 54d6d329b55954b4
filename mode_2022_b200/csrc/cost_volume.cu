// cost_volume.cu -- PSMNet-style concatenation cost volume (reference: models/mode_disparity.py:104-113).
//
// Pure data movement, HBM-write bound: per pair at 1024x512/D=192 the kernel reads 2 x 4.19 MB (L2 resident
// after first touch) and writes 402.65 MB (fp32 NCDHW) or 201.3 MB (bf16 NDHWC).  Every thread produces one
// 128-bit store; stores are fully coalesced in both layouts.  Shift indices are integers -> bit-exact.
#include "common.cuh"
using namespace mode;

// fp32 NCDHW: one thread = 4 consecutive w of one (b, c2, i, h) row.
__global__ void __launch_bounds__(256) cost_volume_f32_kernel(const float* __restrict__ ref, const float* __restrict__ tgt,
                                                              float* __restrict__ cost, int C, int H, int W, int D4, long long total4) {
  const int W4 = W >> 2;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total4; idx += (long long)gridDim.x * blockDim.x) {
    int w = (int)(idx % W4) << 2;
    long long r = idx / W4;
    int h = (int)(r % H);
    r /= H;
    int i = (int)(r % D4);
    r /= D4;
    int c2 = (int)(r % (2 * C));
    int b = (int)(r / (2 * C));
    float4 v;
    if (c2 < C) {
      const float* src = ref + (((size_t)b * C + c2) * H + h) * W + w;
      float4 s = *reinterpret_cast<const float4*>(src);
      v.x = (w + 0 >= i) ? s.x : 0.f;
      v.y = (w + 1 >= i) ? s.y : 0.f;
      v.z = (w + 2 >= i) ? s.z : 0.f;
      v.w = (w + 3 >= i) ? s.w : 0.f;
    } else {
      const float* src = tgt + (((size_t)b * C + (c2 - C)) * H + h) * W;
      v.x = (w + 0 >= i) ? __ldg(src + w + 0 - i) : 0.f;
      v.y = (w + 1 >= i) ? __ldg(src + w + 1 - i) : 0.f;
      v.z = (w + 2 >= i) ? __ldg(src + w + 2 - i) : 0.f;
      v.w = (w + 3 >= i) ? __ldg(src + w + 3 - i) : 0.f;
    }
    st_na_v4(cost + (idx << 2), *reinterpret_cast<uint4*>(&v));
  }
}

// bf16 NDHWC: voxel = 2C bf16; one thread = one 16-byte chunk (8 channels) of one voxel.
__global__ void __launch_bounds__(256) cost_volume_bf16_kernel(const uint16_t* __restrict__ ref, const uint16_t* __restrict__ tgt,
                                                               uint16_t* __restrict__ cost, int C, int H, int W, int D4, long long total8) {
  const int chunks = (2 * C) >> 3, half = C >> 3;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total8; idx += (long long)gridDim.x * blockDim.x) {
    int ch = (int)(idx % chunks);
    long long r = idx / chunks;
    int w = (int)(r % W);
    r /= W;
    int h = (int)(r % H);
    r /= H;
    int i = (int)(r % D4);
    int b = (int)(r / D4);
    uint4 v = make_uint4(0, 0, 0, 0);
    if (w >= i) {
      if (ch < half)
        v = __ldg(reinterpret_cast<const uint4*>(ref + (((size_t)b * H + h) * W + w) * C) + ch);
      else
        v = __ldg(reinterpret_cast<const uint4*>(tgt + (((size_t)b * H + h) * W + (w - i)) * C) + (ch - half));
    }
    st_na_v4(cost + (idx << 3), v);
  }
}

extern "C" int mode_cost_volume_f32(const float* ref, const float* tgt, float* cost, int B, int C, int H, int W, int D4, void* stream) {
  MODE_CHECK_ARG(ref && tgt && cost, "cost_volume_f32: null pointer");
  MODE_CHECK_ARG(B > 0 && C > 0 && H > 0 && W > 0 && D4 > 0, "cost_volume_f32: bad shape B=%d C=%d H=%d W=%d D4=%d", B, C, H, W, D4);
  MODE_CHECK_ARG(W % 4 == 0, "cost_volume_f32: W (%d) must be a multiple of 4", W);
  long long total4 = (long long)B * 2 * C * D4 * H * (W / 4);
  int blocks = (int)std::min<long long>((total4 + 255) / 256, (long long)kNumSMs * 32);
  cost_volume_f32_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(ref, tgt, cost, C, H, W, D4, total4);
  MODE_CHECK_LAUNCH("cost_volume_f32");
  return MODE_OK;
}

extern "C" int mode_cost_volume_16(const mode_h16* ref, const mode_h16* tgt, mode_h16* cost, int B, int C, int H, int W, int D4,
                                     void* stream) {
  MODE_CHECK_ARG(ref && tgt && cost, "cost_volume_16: null pointer");
  MODE_CHECK_ARG(B > 0 && C > 0 && H > 0 && W > 0 && D4 > 0, "cost_volume_16: bad shape B=%d C=%d H=%d W=%d D4=%d", B, C, H, W, D4);
  MODE_CHECK_ARG(C % 8 == 0, "cost_volume_16: C (%d) must be a multiple of 8", C);
  long long total8 = (long long)B * D4 * H * W * (2 * C / 8);
  int blocks = (int)std::min<long long>((total8 + 255) / 256, (long long)kNumSMs * 32);
  cost_volume_bf16_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(ref, tgt, cost, C, H, W, D4, total8);
  MODE_CHECK_LAUNCH("cost_volume_16");
  return MODE_OK;
}
