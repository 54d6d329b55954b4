// api.cu -- error reporting, launch counting, layout helpers.
#include <atomic>
#include <cstdarg>

#include "common.cuh"

namespace mode {
static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }
}  // namespace mode

using namespace mode;

extern "C" int mode_b200_version(void) { return 100; }
extern "C" const char* mode_b200_last_error(void) { return mode::g_err; }
extern "C" unsigned long long mode_b200_launch_count(void) { return mode::g_launches.load(); }

// ---- NCHW fp32 <-> NHWC bf16 -------------------------------------------------------------------
// 32x32 smem transpose tiles: coalesced on both sides.
__global__ void nchw_f32_to_nhwc_bf16_kernel(const float* __restrict__ x, uint16_t* __restrict__ y, int C, int HW, int fmt) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const float* xb = x + (size_t)b * C * HW;
  uint16_t* yb = y + (size_t)b * C * HW;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = c0 + i, p = p0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && p < HW) ? xb[(size_t)c * HW + p] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int p = p0 + i, c = c0 + threadIdx.x;
    if (c < C && p < HW) yb[(size_t)p * C + c] = float_to_h16_bits(tile[threadIdx.x][i], fmt);
  }
}
__global__ void nhwc_bf16_to_nchw_f32_kernel(const uint16_t* __restrict__ x, float* __restrict__ y, int C, int HW, int fmt) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const uint16_t* xb = x + (size_t)b * C * HW;
  float* yb = y + (size_t)b * C * HW;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int p = p0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && p < HW) ? h16_bits_to_float(xb[(size_t)p * C + c], fmt) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = c0 + i, p = p0 + threadIdx.x;
    if (c < C && p < HW) yb[(size_t)c * HW + p] = tile[threadIdx.x][i];
  }
}

extern "C" int mode_nchw_f32_to_nhwc_16(const float* x, mode_h16* y, int B, int C, int HW, int fmt, void* stream) {
  MODE_CHECK_ARG(x && y && B > 0 && C > 0 && HW > 0, "nchw_f32_to_nhwc_bf16: bad arguments");
  dim3 grid(ceil_div(HW, 32), ceil_div(C, 32), B), block(32, 8);
  nchw_f32_to_nhwc_bf16_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(x, y, C, HW, fmt);
  MODE_CHECK_LAUNCH("nchw_f32_to_nhwc_bf16");
  return MODE_OK;
}
extern "C" int mode_nhwc_16_to_nchw_f32(const mode_h16* x, float* y, int B, int C, int HW, int fmt, void* stream) {
  MODE_CHECK_ARG(x && y && B > 0 && C > 0 && HW > 0, "nhwc_bf16_to_nchw_f32: bad arguments");
  dim3 grid(ceil_div(HW, 32), ceil_div(C, 32), B), block(32, 8);
  nhwc_bf16_to_nchw_f32_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(x, y, C, HW, fmt);
  MODE_CHECK_LAUNCH("nhwc_bf16_to_nchw_f32");
  return MODE_OK;
}

// ---- channel concatenation of three NHWC 16-bit maps (the 64 + 128 + 128 channel feature pyramid in front of lastconv,
// reference submodule.py:198: torch.cat((output_raw, output_regular, output_sphere), 1)): one 16-byte chunk per thread,
// 4 chunks in flight.  ATen's generic batched cat reaches ~2.5 TB/s of read+write here; this is a plain streaming copy.
__global__ void __launch_bounds__(256) concat3_nhwc_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, const uint4* __restrict__ c, uint4* __restrict__ out,
                                                           int ka, int kb, int kc, long long total) {
  const int kt = ka + kb + kc;  // 16-byte chunks per output pixel
  for (long long t0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x); t0 < total; t0 += (long long)gridDim.x * blockDim.x * 4) {
    uint4 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long t = t0 + (long long)k * gridDim.x * blockDim.x;
      if (t < total) {
        const long long p = t / kt;
        const int ch = (int)(t - p * kt);
        v[k] = ch < ka ? __ldg(a + p * ka + ch) : (ch < ka + kb ? __ldg(b + p * kb + (ch - ka)) : __ldg(c + p * kc + (ch - ka - kb)));
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long t = t0 + (long long)k * gridDim.x * blockDim.x;
      if (t < total) out[t] = v[k];
    }
  }
}

extern "C" int mode_concat3_nhwc_16(const mode_h16* a, const mode_h16* b, const mode_h16* c, mode_h16* out, long long npix, int Ca, int Cb, int Cc, void* stream) {
  MODE_CHECK_ARG(a && b && c && out && npix > 0, "concat3_nhwc_16: bad arguments");
  MODE_CHECK_ARG(Ca > 0 && Cb > 0 && Cc > 0 && Ca % 8 == 0 && Cb % 8 == 0 && Cc % 8 == 0, "concat3_nhwc_16: channel counts must be multiples of 8");
  const long long total = npix * ((Ca + Cb + Cc) / 8);
  const int blocks = (int)std::min<long long>(ceil_div(total, 256 * 4), (long long)kNumSMs * 16);
  concat3_nhwc_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint4*>(a), reinterpret_cast<const uint4*>(b), reinterpret_cast<const uint4*>(c),
                                                                reinterpret_cast<uint4*>(out), Ca / 8, Cb / 8, Cc / 8, total);
  MODE_CHECK_LAUNCH("concat3_nhwc_16");
  return MODE_OK;
}
