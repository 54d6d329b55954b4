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
